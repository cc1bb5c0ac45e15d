"""fresh process: how long does the first CUDA call of the library take (context creation + module registration)?"""
import ctypes, os, sys, time
t0 = time.perf_counter()
lib = ctypes.CDLL(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tophat_b200", "libtophat_b200.so"))
t1 = time.perf_counter()
p = ctypes.c_void_p(); rc = lib.thb_create(0, ctypes.byref(p)); t2 = time.perf_counter()
q = ctypes.c_void_p(); rc2 = lib.thb_create(0, ctypes.byref(q)); t3 = time.perf_counter()
print("dlopen %.3f s, first thb_create %.3f s (rc %d), second thb_create %.3f s (rc %d), CUDA_MODULE_LOADING=%s" % (
    t1 - t0, t2 - t1, rc, t3 - t2, rc2, os.environ.get("CUDA_MODULE_LOADING", "(default)")), flush=True)
