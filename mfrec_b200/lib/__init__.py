"""Drop-in replacements for the reference's native modules ``mfrec.lib.kmf_train`` and
``mfrec.lib.gd_estimator``: same function names, positional order, defaults, in-place
semantics and exception types, executed by libmfrec_b200 on a B200."""
