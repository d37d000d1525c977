# How long does CUDA initialisation itself take on this box, compared with fq_create?
import ctypes, time, glob, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
t0 = time.perf_counter()
rt = ctypes.CDLL(sorted(glob.glob("/usr/local/cuda/lib64/libcudart.so*"))[0])
t1 = time.perf_counter()
rt.cudaFree(None)
t2 = time.perf_counter()
print(f"dlopen cudart {t1-t0:.3f} s, cudaFree(0) [context init] {t2-t1:.3f} s")
from faqcs_b200.api import Engine, Options
t3 = time.perf_counter()
e = Engine(Options(), device=0)
t4 = time.perf_counter()
print(f"fq_create after init {t4-t3:.3f} s")
e.close() if hasattr(e, "close") else None
