"""dml-b200: the din-mol-Li hot path on B200 behind a C ABI (libdml.so); `dml` is the ctypes mirror of include/dml.h."""
