"""`helpers.utils` as the reference scripts import it (helpers/utils.py).  The label table is read lazily
from metadata/class_labels_indices.csv relative to the CWD, exactly like the reference (helpers/utils.py:38)."""
import csv
import os

from efficientat_b200.helpers.utils import (NAME_TO_WIDTH, exp_rampup, exp_warmup_linear_down,  # noqa: F401
                                            linear_rampdown, mixup)


def _load_labels(path="metadata/class_labels_indices.csv"):
    if not os.path.exists(path):
        return [], []
    with open(path, "r") as f:
        lines = list(csv.reader(f, delimiter=","))
    return [l[2] for l in lines[1:]], [l[1] for l in lines[1:]]


labels, ids = _load_labels()
classes_num = len(labels)
