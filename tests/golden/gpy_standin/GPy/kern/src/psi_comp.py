class PSICOMP_RBF:
    """Psi-statistics helper: constructed by CausalRBF.__init__ (causal_kernels.py:22-24), never called on the path."""


class PSICOMP_RBF_GPU(PSICOMP_RBF):
    pass
