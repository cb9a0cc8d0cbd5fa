"""BASELINE.json configs[2]: level 1 (fast) vs level 9 sweep of deflate on a 4 GiB synthetic corpus, ratio vs reference.
Every level 1..9 is timed device-resident on the whole corpus (text-like and mixed-entropy), every stream is inflated on
the GPU and compared bit-exactly; the reference (oracle restatement of DeflateStream.kt) and libz compress a stated prefix
at levels 1 / 6 / 9, and the engine's size on the same prefix stands beside them.  Writes one JSON document.
usage: level_sweep.py [size_mib=4096] [prefix_mib=256] [out=profiles/level_sweep_r2.json]"""
import json, os, sys, time, zlib
from concurrent.futures import ProcessPoolExecutor
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np

size_mib = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
prefix_mib = int(sys.argv[2]) if len(sys.argv) > 2 else 256
out_path = sys.argv[3] if len(sys.argv) > 3 else os.path.join(ROOT, "profiles", "level_sweep_r2.json")
n, pn = size_mib << 20, min(prefix_mib, size_mib) << 20


def _piece(job):
    import zlibk_b200  # noqa: F401  (puts the package on the path of the worker)
    from zlibk_b200 import corpus
    kind, i, nbytes = job
    if kind == "text":
        return corpus._text_piece(nbytes, corpus.SEED_TEXT + 0x1000 * i)
    return corpus._segment(i % 6, nbytes, corpus.SEED_MIXED + 0x100 * i)


def build(kind):
    """corpus.text_like(n) / corpus.mixed_entropy(n), its pieces generated on every host core"""
    piece = (16 << 20) if kind == "text" else (1 << 20)
    jobs = [(kind, i, min(piece, n - off)) for i, off in enumerate(range(0, n, piece))]
    with ProcessPoolExecutor(max_workers=os.cpu_count()) as ex:
        parts = list(ex.map(_piece, jobs, chunksize=4))
    return np.concatenate(parts)


def cpu_sizes(job):
    from tests.oracle_binding import Oracle
    kind, level, path = job
    data = open(path, "rb").read()
    oracle = Oracle(os.path.join(ROOT, "oracle", "libzk_oracle.so"))
    t0 = time.perf_counter()
    rc, rs, _ = oracle.deflate(data, level)
    dt = time.perf_counter() - t0
    t0 = time.perf_counter()
    z = zlib.compress(data, level)
    dz = time.perf_counter() - t0
    return kind, level, len(rs), len(data) / dt / 1e9, len(z), len(data) / dz / 1e9


def main():
    import torch
    import zlibk_b200

    eng = zlibk_b200.Engine(0)
    dev = torch.device("cuda", 0)
    doc = {"what": "BASELINE.json configs[2]: deflate level sweep on a %d MiB corpus, one B200, device-resident, best of 3; sizes of the "
                   "reference (oracle restatement, one core) and of libz 1.3 on the first %d MiB at levels 1 / 6 / 9" % (size_mib, pn >> 20),
           "size_mib": size_mib, "prefix_mib": pn >> 20, "corpora": {}}
    cpu_jobs = []
    for kind in ("text", "mixed"):
        t0 = time.perf_counter()
        host = build(kind)
        gen_s = time.perf_counter() - t0
        path = "/tmp/level_sweep_prefix_%s.bin" % kind
        host[:pn].tofile(path)
        cpu_jobs += [(kind, level, path) for level in (1, 6, 9)]
        d_in = torch.from_numpy(host).to(dev)
        cap = eng.compress_bound(n)
        d_out = torch.empty(cap, dtype=torch.uint8, device=dev)
        d_back = torch.empty(n, dtype=torch.uint8, device=dev)
        d_res = torch.zeros(4, dtype=torch.int64, device=dev)
        d_res2 = torch.zeros(4, dtype=torch.int64, device=dev)
        d_pref = d_in[:pn]
        rows = {}
        for level in range(1, 10):
            best = 1e9
            for it in range(3):
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                rc = eng.deflate_device(d_in.data_ptr(), n, level, 0, d_out.data_ptr(), cap, d_res.data_ptr())
                eng.synchronize()
                best = min(best, time.perf_counter() - t0)
            m = d_res.cpu().tolist()[0]
            ibest = 1e9
            for it in range(2):
                t0 = time.perf_counter()
                eng.inflate_device(d_out.data_ptr(), m, d_back.data_ptr(), n, d_res2.data_ptr())
                eng.synchronize()
                ibest = min(ibest, time.perf_counter() - t0)
            ok = d_res2.cpu().tolist()[2] == 1 and bool(torch.equal(d_back, d_in))
            assert ok, (kind, level)
            eng.deflate_device(d_pref.data_ptr(), pn, level, 0, d_out.data_ptr(), cap, d_res.data_ptr())
            eng.synchronize()
            rows[level] = {"deflate_gbs": n / best / 1e9, "inflate_gbs": n / ibest / 1e9, "size": m, "ratio": m / n,
                           "gpu_inflate_bit_exact": ok, "size_on_prefix": d_res.cpu().tolist()[0]}
            print(kind, level, "%.1f GB/s" % (n / best / 1e9), "%.4f" % (m / n), flush=True)
        doc["corpora"][kind] = {"bytes": n, "generated_in_s": gen_s, "levels": rows}
        del d_in, d_out, d_back, host
        torch.cuda.empty_cache()
    with ProcessPoolExecutor(max_workers=6) as ex:
        for kind, level, ref_size, ref_gbs, z_size, z_gbs in ex.map(cpu_sizes, cpu_jobs):
            row = doc["corpora"][kind]["levels"][level]
            row.update({"reference_size_on_prefix": ref_size, "reference_gbs_one_core": ref_gbs, "libz_size_on_prefix": z_size,
                        "libz_gbs_one_core": z_gbs, "size_vs_reference": row["size_on_prefix"] / ref_size,
                        "size_vs_libz": row["size_on_prefix"] / z_size})
    for kind in doc["corpora"]:
        lv = doc["corpora"][kind]["levels"]
        doc["corpora"][kind]["checks"] = {
            "level9_le_level6_le_level1": lv[9]["size"] <= lv[6]["size"] <= lv[1]["size"],
            "within_1.02_of_reference_at_1_6_9": all(lv[l]["size_vs_reference"] <= 1.02 for l in (1, 6, 9)),
            "within_1.02_of_libz_at_6_9": all(lv[l]["size_vs_libz"] <= 1.02 for l in (6, 9)),
            "level9_over_level1_speed": lv[9]["deflate_gbs"] / lv[1]["deflate_gbs"]}
    json.dump(doc, open(out_path, "w"), indent=1)
    print(json.dumps({k: v["checks"] for k, v in doc["corpora"].items()}))


if __name__ == "__main__":
    main()
