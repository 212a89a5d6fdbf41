"""Mirror of ``/root/reference/core/transformer``: only the op seam (``attention``) and parameter containers
of the decode path live here; the arithmetic runs in ``edgerunner_b200/csrc``."""
