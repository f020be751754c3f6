"""sparkfm_b200 -- B200-native (sm_100a) FM scorer + mini-batch SGD trainer behind SparkFM's
learner-plugin interface.  The product is libsparkfm_b200.so (C ABI: include/sparkfm_b200.h);
this package is the ctypes binding plus a Python mirror of the reference's operator API."""
from . import _lib  # noqa: F401
from .api import (ALS, DataSet, FM, FMLearn, FMModel, FMUtils, FMWithSGD, FactorizationMachines,  # noqa: F401
                  LabeledPoint, Model, SGD, SparseVector, Task)
from .handle import (Handle, device_count, format_libfm, pack_onehot, parse_libfm,  # noqa: F401
                     partition_rows, sample_rows)

__all__ = ["ALS", "DataSet", "FM", "FMLearn", "FMModel", "FMUtils", "FMWithSGD", "FactorizationMachines",
           "LabeledPoint", "Model", "SGD", "SparseVector", "Task", "Handle", "device_count",
           "format_libfm", "pack_onehot", "parse_libfm", "partition_rows", "sample_rows"]
