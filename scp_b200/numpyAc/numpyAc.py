"""``numpyAc.arithmeticCoding`` with the reference's interface (numpyAc/numpyAc.py:116-137): same arguments, same
assertions / ValueErrors, byte-identical output.  The PMF -> uint16 CDF conversion (:80-114) runs in the
scp_pmf_to_cdf CUDA kernel, the 32-bit range coder (numpyAc_backend.cpp:245-323) in scp_range_encode."""
import numpy as np
import torch

from .. import coder

PRECISION = 16


def pdf_convert_to_cdf_and_normalize(pdf):
    """numpyAc.py:109-114 (host helper kept for callers that want the float CDF)."""
    assert pdf.ndim == 2
    cdfF = np.cumsum(pdf, axis=1)
    cdfF = cdfF / cdfF[:, -1:]
    return np.hstack((np.zeros((pdf.shape[0], 1)), cdfF))


def _convert_to_int_and_normalize(cdf_float, needs_normalization):
    """numpyAc.py:80-107"""
    Lp = cdf_float.shape[-1]
    new_max_value = 2 ** PRECISION - ((Lp - 1) if needs_normalization else 0)
    cdf = np.round(cdf_float * new_max_value).astype(np.int16)
    if needs_normalization:
        cdf += np.arange(Lp).astype(np.int16)
    return cdf


class arithmeticCoding():
    def __init__(self) -> None:
        self.binfile = None
        self.sysNum = None
        self.byte_stream = None

    def encode(self, pdf, sym, binfile=None):
        assert pdf.shape[0] == sym.shape[0]
        assert pdf.ndim == 2 and sym.ndim == 1
        if sym.dtype != np.int16:
            raise ValueError('Symbols must be int16!')                      # numpyAc.py:57-58
        if pdf.shape[1] != 255:
            raise ValueError('the CUDA CDF kernel is specialised for the 255 occupancy symbols of SCP')
        self.sysNum = sym.shape[0]
        pdf = np.ascontiguousarray(pdf, dtype=np.float32)
        if self.sysNum:
            if pdf.min() < 0:                                              # numpyAc.py:32-39 (check_input_bounds)
                raise ValueError(f'cdf_float.min() == {pdf.min()}, should be >=0.!')
            if sym.max() >= 255 or sym.min() < 0:
                raise ValueError(f'sym.max() == {sym.max()}, should be <=Lp - 1.!')
        d = coder.pmf_to_cdf(torch.from_numpy(pdf).cuda(), sym=torch.from_numpy(np.ascontiguousarray(sym)).cuda(),
                             is_logits=False, want_interval=True)
        self.byte_stream = coder.range_encode(d["interval"].cpu().numpy())
        real_bits = len(self.byte_stream) * 8
        if binfile is not None:
            with open(binfile, 'wb') as fout:
                fout.write(self.byte_stream)
        return self.byte_stream, real_bits
