"""Drop-in for the reference's ``numpyAc`` package (numpyAc/numpyAc.py:116-137): ``arithmeticCoding().encode``."""
from .numpyAc import arithmeticCoding, pdf_convert_to_cdf_and_normalize, _convert_to_int_and_normalize  # noqa: F401
