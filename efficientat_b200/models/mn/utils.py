"""Shape arithmetic shared by the MN / DyMN builders (reference models/mn/utils.py:8-26)."""
import math


def make_divisible(v, divisor, min_value=None):
    """Round `v` to a multiple of `divisor`, never dropping more than 10 % (mn/utils.py:8-21)."""
    floor = divisor if min_value is None else min_value
    rounded = max(floor, int(v + divisor / 2) // divisor * divisor)
    return rounded + divisor if rounded < 0.9 * v else rounded


def cnn_out_size(in_size, padding, dilation, kernel, stride):
    """mn/utils.py:24-26"""
    return math.floor((in_size + 2 * padding - dilation * (kernel - 1) - 1) / stride + 1)
