"""TensorFlow-1.2 / GPflow-0.5 API shim (torch fp64 backend) used ONLY to execute the unmodified
reference files under /root/reference and freeze their outputs as golden vectors
(oracle/make_golden.py -> tests/golden/).  TEST INFRASTRUCTURE ONLY; never imported by the product.

The reference's own arithmetic (every line of matern12_spectral_mixture.py, sgpr_ss.py, pdgp.py,
likelihoods.py, methods.py, window_overlap.py) runs verbatim; only the TF/GPflow *library* calls it
makes are served by this shim.  Python-2 semantics the files rely on (integer '/', builtin
``reduce``, implicit relative imports) are restored by an AST pass in loader.py.
"""
