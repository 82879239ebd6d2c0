"""Drop-in for the reference's ``decode_ehem_mullevel.py`` (three sub-octrees in one stream, :179-206)."""
from . import decode_ehem as _dec

extract_info = _dec.extract_info


def decodeOct(binfile, oct_data_seqs, model, context_size=8192, anc_k=4):
    """decode_ehem_mullevel.py:179-206; ``oct_data_seqs`` = the three original sequences (or None)."""
    import numpy as np
    seq = None if oct_data_seqs is None else np.concatenate([np.asarray(s).reshape(-1) for s in oct_data_seqs])
    return _dec.decodeOct(binfile, seq, model, context_size, anc_k, mullevel=True)


def main(args):
    """decode_ehem_mullevel.py:209-274"""
    return _dec.main(args, mullevel=True)


get_args = _dec.get_args

if __name__ == "__main__":
    main(get_args())
