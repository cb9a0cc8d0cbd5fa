"""Host-buffer deflate / inflate from CUDA-pinned caller memory against ordinary pageable memory (what a Kotlin
ByteArray.usePinned hands over: pinned for the GC, pageable for CUDA) and against memory registered once with
zk_host_register.  usage: pageable_e2e.py [mib=1024]"""
import ctypes, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import zlibk_b200
from zlibk_b200 import corpus

mib = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
n = mib << 20
eng = zlibk_b200.Engine(0)
lib = zlibk_b200.library()
host = corpus.mixed_entropy(n)
cap = eng.compress_bound(n)


def timed(fn, reps=3):
    best = 1e9
    for it in range(reps + 1):
        t0 = time.perf_counter()
        r = fn()
        dt = time.perf_counter() - t0
        if it:
            best = min(best, dt)
    return best, r


def run(kind):
    if kind == "pinned":
        h_in = torch.from_numpy(host).pin_memory().numpy()
        h_out = torch.empty(cap, dtype=torch.uint8).pin_memory().numpy()
        h_back = torch.empty(n, dtype=torch.uint8).pin_memory().numpy()
    else:
        h_in = host.copy()
        h_out = np.empty(cap, dtype=np.uint8)
        h_back = np.empty(n, dtype=np.uint8)
        h_out[:] = 0
        h_back[:] = 0  # touch: page faults are not what is measured
    reg = []
    eng.set_option("auto_register", 1 if kind == "auto" else 0)
    if kind == "registered":
        t0 = time.perf_counter()
        for a in (h_in, h_out, h_back):
            rc = lib.zk_host_register(a.ctypes.data, a.nbytes)
            assert rc == 0, rc
            reg.append(a)
        print("  zk_host_register of %.1f GiB: %.1f ms (once per buffer, not per call)" % ((n * 2 + cap) / 2**30, (time.perf_counter() - t0) * 1e3))
    dt, (rc, s, _) = timed(lambda: eng.deflate(h_in, 6, out=h_out))
    assert rc == 0
    m = len(s)
    stream = h_out[:m]
    dti, (rci, back, used) = timed(lambda: eng.inflate(stream, n, out=h_back))
    assert rci == 1 and used == m and np.array_equal(h_back, host)
    print("%-10s deflate %.2f ms %.1f GB/s | inflate %.2f ms %.1f GB/s" % (kind, dt * 1e3, n / dt / 1e9, dti * 1e3, n / dti / 1e9), flush=True)
    for a in reg:
        lib.zk_host_unregister(a.ctypes.data)


for kind in ("pinned", "pageable", "registered", "auto"):
    run(kind)
