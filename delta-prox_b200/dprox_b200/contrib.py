"""Application glue of the reference that sits directly on the hot path (dprox/contrib): the CS-MRI solver.

`CustomADMM` is contrib/csmri.py:156-171: ADMM with the prox evaluated FIRST (x = prox(z - u); z = solve(x + u);
u += x - z) for a single deep prior and the `csmri` closed-form data term, on complex iterates; state `(x, [z], [u])`.
It runs on the generic engine: native glue kernels, the denoiser as an external prox, `dpx_csmri_prox` as the x-update.
"""
from .algo import ADMM


class CustomADMM(ADMM):
    method = "custom_admm"
