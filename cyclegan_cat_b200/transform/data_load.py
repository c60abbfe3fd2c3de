"""`normalize` of the reference input pipeline (transform/data_load.py:31-34), the only
piece of it on the predict.py boundary (predict.py:9,23).  The tf.data TFRecord pipeline
itself is out of scope."""
import numpy as np


def normalize(tensor) -> np.ndarray:
    image = np.asarray(tensor, dtype=np.float32)
    return (image / np.float32(127.5)) - np.float32(1.0)
