from efficientat_b200.models.mn.utils import cnn_out_size, make_divisible  # noqa: F401
