"""GPU: long-row scan timings — the single-pass look-back kernel against the three-pass chunked path
(PDLB200_SCAN=3pass in a child process), several types and sizes.  Prints one JSON line per case."""
import json
import os
import subprocess
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))


def main():
    import torch
    from pdl_b200 import trans as P, types as T
    from pdl_b200.engine import CudaEngine
    sys.path.insert(0, str(Path(__file__).resolve().parent))
    from microbench import wrap, timeit, PEAK
    eng = CudaEngine()
    dev = "cuda"
    g = torch.Generator(device=dev).manual_seed(3)
    mode = os.environ.get("PDLB200_SCAN", "onepass")
    cases = [("float", T.F, torch.float32, 2**28, 1), ("float", T.F, torch.float32, 2**24, 1),
             ("float", T.F, torch.float32, 2**22, 64), ("double", T.D, torch.float64, 2**27, 1),
             ("long", T.L, torch.int32, 2**28, 1), ("longlong", T.LL, torch.int64, 2**27, 1)]
    for name, t, dt, n, rows in cases:
        x = torch.randint(-8, 9, (n * rows,), device=dev, generator=g).to(dt)
        px = wrap(eng, x, t, [n, rows] if rows > 1 else [n])
        out = P.PDL.empty(t, px.dims, eng)
        for bad in (False, True):
            px.badflag = bad
            c0 = eng.launch_count()
            P.run_op("cumusumover", [px], [out])
            nl = eng.launch_count() - c0
            ms = timeit(lambda: P.run_op("cumusumover", [px], [out]), 20)
            by = 2 * x.element_size() * n * rows
            print(json.dumps({"mode": mode, "case": f"cumusumover {name}[{n}{',' + str(rows) if rows > 1 else ''}]",
                              "bad": bad, "launches": nl, "ms": round(ms, 4), "gbs": round(by / ms / 1e6, 1),
                              "frac": round(by / ms / 1e6 / PEAK, 3)}), flush=True)
        del x, px, out


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "both":
        for m in ("onepass", "3pass"):
            subprocess.run([sys.executable, __file__], env=dict(os.environ, PDLB200_SCAN=m), check=False)
    else:
        main()
