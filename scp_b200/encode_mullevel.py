"""Drop-in for the reference's ``encode_mullevel.py`` (three sub-octrees per frame, README.md:80)."""
from . import encode as _enc


def compress_ehem(batch, outputfile, model, args):
    """encode_mullevel.py:88-157"""
    return _enc.compress_ehem(batch, outputfile, model, args, mullevel=True)


compress = _enc.compress_mullevel      # encode_mullevel.py:23-85
get_args = _enc.get_args


def main(args):
    return _enc.main(args, mullevel=True)


if __name__ == "__main__":
    main(get_args())
