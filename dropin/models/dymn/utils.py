from efficientat_b200.models.dymn.utils import cnn_out_size, make_divisible  # noqa: F401
