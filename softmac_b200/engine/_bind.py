import ctypes as C
ip = C.POINTER(C.c_int32)
