"""FFI namespace package (the reference builds a cffi extension here; this one is ctypes)."""
