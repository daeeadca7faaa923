"""B200-native AP-VAST sound-zone filter engine (drop-in for ``Python/apvast.py`` of
macoustics/ap-vast-unofficial).  ``from ap_vast_unofficial_b200 import apvast, jdiag``."""
from .apvast import apvast, jdiag, EXPERIMENTAL_NORMALIZE_GAINS, EXPERIMENTAL_REGULARIZATION  # noqa: F401

__all__ = ["apvast", "jdiag"]
