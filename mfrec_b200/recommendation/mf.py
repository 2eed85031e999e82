"""``MFRecommender``: shared pieces of the factor-model recommenders
(reference: mfrec/recommendation/mf.py:33-194)."""
import numpy as np

from mfrec_b200 import _native
from mfrec_b200.recommendation.base import BaseRecommender


class MFRecommender(BaseRecommender):
    # predictor method name -> predictor of the C ABI (include/mfrec_b200.h MFREC_PRED_*)
    NATIVE_PREDICTORS = {}

    def __init__(self, nbr_users=4, nbr_items=6, parameters=False):
        BaseRecommender.__init__(self, nbr_users, nbr_items, parameters)
        self.neighborhood = 500

    def clamping(self, value, min=1.0, max=5.0):
        # the bounds are hard-coded in the reference too (mf.py:44-53)
        if value > 5.0:
            value = 5.0
        if value < 1.0:
            value = 1.0
        return value

    def warmyup(self):
        self.relationship_matrix_csc = self.relationship_matrix.tocsc()

    def predict_rating_by_label(self, user_label, item_label, predictor='predict_logistic'):
        try:
            item_index = self.items_index[item_label]
            user_index = self.users_index[user_label]
            return getattr(self, predictor)(item_index, user_index)
        except KeyError:
            return self.overall_avg

    # ---- BellKor bias statistics (mf.py:78-121): one device pass instead of per-row scipy slicing
    def _bias_stats(self):
        idx, r = self.get_ratings(randomize_order=False)
        mu, ib, ub = _native.bias_stats(idx, r, self.nbr_items, self.nbr_users,
                                        getattr(self, 'K2', 0.01), getattr(self, 'K3', 0.01))
        return mu, ib, ub

    def compute_items_bias_bk(self):
        if not self.overall_bias:
            self.compute_overall_avg()
        _, self.items_bias, _ = self._bias_stats()
        self.relationship_matrix_csc = self.relationship_matrix.tocsc()

    def compute_users_bias_bk(self):
        if not self.overall_bias:
            self.compute_overall_avg()
        _, ib, self.users_bias = self._bias_stats()
        if self.items_bias is None:
            self.items_bias = ib
        self.relationship_matrix_csr = self.relationship_matrix.tocsr()

    # ---- initialisation (mf.py:124-141) --------------------------------------------------------
    def init_feature_normal(self, mean=0.0, std=0.1):
        self.svd_u = np.random.normal(mean, std, [self.dimensionality, self.nbr_items])
        self.svd_v = np.random.normal(mean, std, [self.dimensionality, self.nbr_users])

    def init_user_features(self, user_index, mean=0.0, std=0.1):
        self.svd_v[:, user_index] = np.random.normal(mean, std, self.dimensionality)

    def init_item_features(self, item_index, mean=0.0, std=0.1):
        self.svd_u[:, item_index] = np.random.normal(mean, std, self.dimensionality)

    # ---- top-N (mf.py:144-193) ---------------------------------------------------------------------
    def _rated_items(self, user_index):
        """Row ids the reference reads from ``relationship_matrix_csc[:, user_index]`` (the matrix
        the subclass stored there; KMF keeps the transpose, so these are the user's items)."""
        col = self.relationship_matrix_csc[:, user_index].tocoo()
        return np.sort(col.row[col.data != 0]).astype(np.int32)

    def _topn(self, user_index, n_candidates, nbr_recommendations, predictor):
        rated = self._rated_items(user_index)
        indptr = np.array([0, rated.shape[0]], dtype=np.int64)
        items, scores, counts = _native.topn(
            self.NATIVE_PREDICTORS[predictor], self.svd_u, self.svd_v,
            np.array([user_index], dtype=np.int32), n_candidates, indptr, rated,
            nbr_recommendations, self.overall_bias or 0.0, self.items_bias, self.users_bias,
            self.min_rating, self.max_rating)
        c = int(counts[0])
        return [int(i) for i in items[0, :c]], [float(s) for s in scores[0, :c]]

    def find_recommended_items(self, user_index=None, user_label=None, nbr_recommendations=5,
                               output_label=False, predictor='predict'):
        """Scores items for one user, drops the rated ones, returns the best N.  Reference quirks
        kept: the sampled item ids are drawn (the RNG advances) but the loop scores
        the enumeration index, i.e. items ``0 .. neighborhood-1``; the user's own index is added to
        the excluded item ids; exact zeros and NaNs are dropped."""
        if user_index is None:
            user_index = self.users_index[user_label]
        self.neighborhood = min([self.neighborhood, self.nbr_items])
        self.get_items_subset(count=self.neighborhood)
        name = predictor if predictor in self.NATIVE_PREDICTORS else self._predict_alias(predictor)
        items, scores = self._topn(user_index, self.neighborhood, nbr_recommendations, name)
        if output_label:
            items = [self.items_label[i] for i in items]
        return items, scores

    # ---- item-item similarity in factor space (base.py:1420-1466) ----------------------------------
    def _similar_rows(self, rows, query, nbr_recommendations, similarity_threshold, similarities_output,
                      method, transform=None):
        """Top neighbours of row ``query`` of ``rows`` ([n, d]) by cosine / Pearson similarity.
        Both are dot products of normalised rows, i.e. one row of a Gram matrix: scored and ranked
        on the device by the top-N kernel (the query itself is excluded there -- the reference
        drops the first entry of the sorted list, which is the item itself)."""
        if method not in ('cosine', 'pearson'):
            raise NotImplementedError("similarity method %r: only 'cosine' and 'pearson' run on the device" % method)
        x = np.array(rows, dtype=np.float64)
        if method == 'pearson':
            x = x - x.mean(axis=1, keepdims=True)
        norm = np.sqrt((x * x).sum(axis=1, keepdims=True))
        x = np.divide(x, norm, out=np.zeros_like(x), where=norm > 0)
        xt = np.ascontiguousarray(x.T)                      # [d, n]: "item factors" and "user factors"
        n = x.shape[0]
        N = n - 1 if nbr_recommendations == 'All' else int(nbr_recommendations)
        items, scores, counts = _native.topn('predict_dot', xt, xt, np.array([query], dtype=np.int32), n,
                                             None, None, max(N, 1))
        c = int(counts[0])
        ids = [int(i) for i in items[0, :c]]
        sims = [float(v) for v in scores[0, :c]]
        if transform is not None:
            sims = [transform(v) for v in sims]
        if similarity_threshold:
            keep = [j for j, v in enumerate(sims) if v > similarity_threshold]
            ids, sims = [ids[j] for j in keep], [sims[j] for j in keep]
        ids, sims = ids[:N], sims[:N]
        return (ids, sims) if similarities_output else ids

    def similar_items(self, item_index, nbr_recommendations=2, similarity_threshold=False,
                      similarities_output=False, method='cosine'):
        return self._similar_rows(self.svd_u.T, item_index, nbr_recommendations, similarity_threshold,
                                  similarities_output, method)

    def similar_items_by_label(self, item_label, nbr_recommendations=2, similarity_threshold=False,
                               similarities_output=False, method='cosine'):
        out = self.similar_items(self.items_index[item_label], nbr_recommendations, similarity_threshold,
                                 True, method)
        labels = [self.items_label[i] for i in out[0]]
        return (labels, out[1]) if similarities_output else labels

    def _predict_alias(self, predictor):
        """'predict' is a class attribute aliasing one of the named predictors."""
        target = getattr(type(self), predictor)
        for name in self.NATIVE_PREDICTORS:
            if getattr(type(self), name, None) is target:
                return name
        raise KeyError(predictor)
