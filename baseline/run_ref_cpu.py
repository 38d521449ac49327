"""Worker of bench.py's CPU legs: times the reference's own CPU implementation of the path on ONE (batch, head) slice
[1,1,S,D] fp32 causal -- `aule._cpu_attention` of the UNMODIFIED reference package installed under baseline/_ref
(python/aule/__init__.py:247-271; `pip install --no-index --no-deps --target baseline/_ref <copy of /root/reference/python>`,
see DESIGN.md), or, when that install is absent, the oracle's restatement of it (oracle/attention_oracle.py::cpu_attention).
Runs in its own process so that the reference's `aule` package never shares an interpreter with this repository's `aule`.
usage: python baseline/run_ref_cpu.py S D nslices seed   -> one JSON line {kind, seconds, nslices, checksum}"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "baseline", "_ref")


def load():
    if os.path.isdir(os.path.join(REF, "aule")):
        sys.path.insert(0, REF)
        import aule                                  # the reference package; no libaule.so beside it -> NumPy path only
        return "reference", aule._cpu_attention
    sys.path.insert(0, ROOT)
    from oracle.attention_oracle import cpu_attention
    return "port", cpu_attention


def main():
    S, D, n, seed = (int(x) for x in sys.argv[1:5])
    import numpy as np
    kind, fn = load()
    rng = np.random.RandomState(seed)
    q, k, v = (rng.randn(1, 1, S, D).astype(np.float32) for _ in range(3))
    chk = 0.0
    t0 = time.perf_counter()
    for _ in range(n):
        o = fn(q, k, v, causal=True)
        chk += float(o[0, 0, -1, 0])
    dt = time.perf_counter() - t0
    print(json.dumps({"kind": kind, "seconds": dt, "nslices": n, "checksum": chk}))


if __name__ == "__main__":
    main()
