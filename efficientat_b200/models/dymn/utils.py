"""reference models/dymn/utils.py duplicates the MN helpers; re-export them."""
from ..mn.utils import cnn_out_size, make_divisible  # noqa: F401
