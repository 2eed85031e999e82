"""``BaseRecommender``: the rating store and label bookkeeping the trainers sit on
(reference: mfrec/recommendation/base.py:68-286, 504-508, 797-853, 1043-1146).

Only the part of the reference class that the SGD path touches is restated: the scipy
``lil_matrix`` of ratings, label <-> index maps, parameter plumbing, COO extraction
(``get_ratings``) and the overall average.  Persistence to MongoDB / Neo4j, the kNN and
similarity helpers are outside the hot path."""
import logging

import numpy as np
from scipy.sparse import coo_matrix, find, lil_matrix


class Error(Exception):
    """The reference raises ``Error`` for bad parameters (base.py:199); the name is kept."""


class BaseRecommender(object):
    PARAMETERS_INDEX = {}
    _logger_name = 'mfrec.recommender'

    def __init__(self, nbr_users=4, nbr_items=6, parameters=None):
        self.logger = logging.getLogger(self._logger_name)
        self.dimensionality = 40
        self.min_rating = 1.0
        self.max_rating = 5.0
        self.relationship_matrix = None
        self.relationship_matrix_csc = None
        self.relationship_matrix_csr = None
        self.items_index = {}
        self.items_label = []
        self.users_index = {}
        self.users_label = []
        self.svd_u = None       # ITEM factors [k, nbr_items]  (base.py:144-146)
        self.svd_v = None       # USER factors [k, nbr_users]
        self.users_bias = None
        self.items_bias = None
        self.overall_bias = None
        self.metadata = {}
        if parameters:
            self.set_parameters(parameters)
        self.initialize_relationship_matrix(int(nbr_users), int(nbr_items))

    # ---- sizes ---------------------------------------------------------------------------
    @property
    def nbr_users(self):
        return len(self.users_label)

    @property
    def nbr_items(self):
        return len(self.items_label)

    @property
    def overall_avg(self):
        return self.overall_bias

    # ---- parameters (base.py:180-199) --------------------------------------------------------
    def set_parameters(self, parameters):
        for key, value in parameters.items():
            try:
                setattr(self, self.PARAMETERS_INDEX[key], value)
            except KeyError:
                raise Error('Wrong parameters')

    def set_dimensionality(self, new_dim_value):
        self.dimensionality = new_dim_value

    # ---- rating store (base.py:266-286, 815-836) ------------------------------------------------
    def initialize_relationship_matrix(self, nbr_users, nbr_items):
        self.relationship_matrix = lil_matrix((nbr_users, nbr_items))
        self.items_label = ['item' + str(i) for i in range(nbr_items)]
        self.items_index = dict((lab, i) for i, lab in enumerate(self.items_label))
        self.users_label = ['user' + str(u) for u in range(nbr_users)]
        self.users_index = dict((lab, u) for u, lab in enumerate(self.users_label))

    def set_item(self, user, items_list):
        for item in items_list:
            self.relationship_matrix[int(self.users_index[user]),
                                     int(self.items_index[item['label']])] = float(item['value'])

    def set_item_by_id(self, user_index, item_index, value):
        self.relationship_matrix[user_index, item_index] = float(value)

    def set_item_by_label(self, user, item, value):
        self.relationship_matrix[int(self.users_index[user]), int(self.items_index[item])] = float(value)

    def set_ratings(self, ratings_index, ratings):
        """Bulk ingestion (not in the reference, which takes one ``set_item_by_id`` call per
        rating): the same matrix as ``set_item_by_id`` over every row -- later duplicates win,
        zeros mean "absent" -- built in one shot."""
        ratings_index = np.asarray(ratings_index)
        ratings = np.asarray(ratings, dtype=np.float64)
        shape = self.relationship_matrix.shape
        key = ratings_index[:, 0].astype(np.int64) * shape[1] + ratings_index[:, 1]
        # keep the LAST occurrence of a duplicated (user, item)
        _, first_rev = np.unique(key[::-1], return_index=True)
        keep = key.shape[0] - 1 - first_rev
        m = coo_matrix((ratings[keep], (ratings_index[keep, 0], ratings_index[keep, 1])), shape=shape)
        self.relationship_matrix = m.tolil()

    def set_user_label(self, user_index, label):
        del self.users_index[self.users_label[user_index]]
        self.users_index[label] = user_index
        self.users_label[user_index] = label

    def set_item_label(self, item_index, label):
        del self.items_index[self.items_label[item_index]]
        self.items_index[label] = item_index
        self.items_label[item_index] = label

    def build_index(self):
        self.users_index = dict((lab, i) for i, lab in enumerate(self.users_label))
        self.items_index = dict((lab, i) for i, lab in enumerate(self.items_label))

    def ratings_iterator(self):
        cx = self.relationship_matrix.tocoo()
        return zip(cx.row, cx.col, cx.data)

    def get_ratings(self, randomize_order=False):
        """COO extraction (base.py:1115-1131): rows in scipy's lil -> coo order (by user, then
        ascending item), ``ratings_index int32 [nnz, 2] = (user, item)``, ``ratings float64``;
        ``randomize_order`` shuffles ONCE with the global legacy numpy RNG.  Vectorised, but the
        arrays and the RNG consumption are identical to the reference's per-rating loop."""
        cx = self.relationship_matrix.tocoo()
        keep = cx.data != 0          # find() drops explicit zeros (base.py:1119)
        ratings = np.ascontiguousarray(cx.data[keep], dtype=np.float64)
        ratings_index = np.empty((ratings.shape[0], 2), dtype=np.int32)
        ratings_index[:, 0] = cx.row[keep]
        ratings_index[:, 1] = cx.col[keep]
        index = np.arange(ratings.shape[0])
        if randomize_order:
            np.random.shuffle(index)
        return ratings_index[index], ratings[index]

    def get_items_subset(self, count=100, method='random'):
        ids = np.arange(self.nbr_items)
        np.random.shuffle(ids)
        return ids[0:count]

    # ---- statistics (base.py:504-508) ------------------------------------------------------------
    def compute_overall_avg(self):
        self.overall_bias = find(self.relationship_matrix)[2].astype(float).mean()

    # ---- model snapshot (base.py:805-812) -------------------------------------------------------
    def save_model_snapshot(self, filename):
        np.savez(filename + '_model_snapshot.npz', svd_u=self.svd_u, svd_v=self.svd_v)

    def load_model_snapshot(self, filename):
        svd = np.load(filename + '_model_snapshot.npz')
        self.svd_u = svd['svd_u']
        self.svd_v = svd['svd_v']

    def _get_new_item_id(self):
        new_id = len(self.items_label)
        self.items_label.append('item' + str(new_id))
        return new_id

    def _get_new_user_id(self):
        new_id = len(self.users_label)
        self.users_label.append('user' + str(new_id))
        return new_id
