"""Generates the golden fixtures of tests/golden/ from the oracle (the reference is Rust and cannot run in this
image, so these are ORACLE outputs, not reference outputs: "parity unpinned" for the transcript-dependent bytes,
see DESIGN.md section 4):

  e2e_proofs.json    the two reference e2e configurations (tests/e2e_goldilocks.rs, tests/e2e_babybear.rs) proved by
                     the pure-Python restatement oracle/pyref.py, and by the C restatement (must agree byte for byte)
  scale_proofs.json  sha256 + length of whole proofs at BASELINE sizes (2^14 .. 2^22 rows, both fields, blowup 4 / 8,
                     binary and 4-ary trees, config 3a and the headline shape) proved by the C restatement
                     oracle/liboracle.so (or_stark_prove) on the synthetic AIR of ministark_b200/synth.py

  python tests/golden/make_golden.py [--force] [--skip-big]
"""
import hashlib
import json
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
from ministark_b200 import synth  # noqa: E402
from oracle import oracle as O  # noqa: E402
from oracle import pyref as R  # noqa: E402

O.build()
force = "--force" in sys.argv
skip_big = "--skip-big" in sys.argv
threads = os.cpu_count() or 1

# ---- the reference's e2e configurations --------------------------------------------------------------
out = {}
for F, steps in ((R.Goldilocks, 9), (R.BabyBear, 7)):
    claim = R.FibonacciClaim(F, steps)
    trace = claim.trace(2)
    cfg = R.StarkConfig(F, 20, 2, trace.step_number(), trace.constrain_number())
    proof = R.Stark(cfg).prove(claim, 2)
    raw = R.serialize_proof(F, proof)
    tr = np.array(trace.data, dtype=np.uint64).reshape(trace.length, trace.width)
    c_raw = O.stark_prove(F.field_id, 20, 2, steps, trace.constrain_number(), tr, np.array(trace.linear_rows, dtype=np.uint64)).tobytes()
    assert c_raw == raw, "C and Python restatements disagree"
    out[F.name] = {
        "steps": steps, "security_bits": 20, "blowup": 2, "rounds": cfg.rounds,
        "constrain_queries": cfg.constrain_queries, "fri_queries": cfg.fri_config.queries,
        "padding_value": str(trace.data[-1]),
        "trace_commit": proof.trace_commit.hex(),
        "constrain_trace_commit": proof.constrain_trace_commit.hex(),
        "arthur": proof.arthur.hex(),
        "proof_len": len(raw),
        "proof_sha256": hashlib.sha256(raw).hexdigest(),
    }
with open(os.path.join(HERE, "e2e_proofs.json"), "w") as fh:
    json.dump(out, fh, indent=1, sort_keys=True)
print(json.dumps({k: v["proof_sha256"] for k, v in out.items()}, indent=1))

# ---- whole proofs at BASELINE sizes ---------------------------------------------------------------------
# (name, field, log2 rows, W (T = W, C = 2W), blowup, security bits, inner_children, trace seed)
SEED = 0x5EED000000000000
SHAPES = [
    ("gl_2^14x8_b4", 0, 14, 4, 4, 100, 2, SEED),
    ("bb_2^14x8_b4", 1, 14, 4, 4, 100, 2, SEED),
    ("gl_2^16x16_b8", 0, 16, 8, 8, 100, 2, SEED),
    ("bb_2^16x16_b8", 1, 16, 8, 8, 100, 2, SEED),
    ("gl_2^15x8_b8_4ary", 0, 15, 4, 8, 100, 4, SEED),
    ("bb_2^15x8_b8_4ary", 1, 15, 4, 8, 100, 4, SEED),
    ("gl_2^16x64_b4", 0, 16, 32, 4, 100, 2, SEED),
    ("config3a_gl_2^20x16_b8", 0, 20, 8, 8, 100, 2, SEED + 3),
    ("headline_gl_2^22x32_b4", 0, 22, 16, 4, 100, 2, SEED + 1),
]
path = os.path.join(HERE, "scale_proofs.json")
scale = {}
if os.path.exists(path) and not force:
    with open(path) as fh:
        scale = json.load(fh)
for name, field, logn, w, B, sec, k, seed in SHAPES:
    if name in scale or (skip_big and logn >= 20):
        continue
    n = 1 << logn
    tr = synth.synth_trace(field, n, w, seed=seed).astype(np.uint64)
    mat = synth.synth_linear_matrix(field, n, w).astype(np.uint64)
    t0 = time.time()
    raw, ms = O.stark_prove(field, sec, B, n - 1, 2 * w, tr, mat, inner_children=k, threads=threads, want_timings=True)
    rounds, cq, fq = O.stark_derive(field, sec, B, n - 1)
    scale[name] = {
        "field": field, "log_rows": logn, "w": w, "t": w, "blowup": B, "security_bits": sec, "inner_children": k,
        "trace_seed": seed, "rounds": rounds, "constrain_queries": cq, "fri_queries": fq,
        "proof_len": int(raw.size), "proof_sha256": hashlib.sha256(raw.tobytes()).hexdigest(),
        "trace_commit": raw[8 + 8 + 8 + 64 * rounds:][:32].tobytes().hex(),
        "oracle_wall_s": round(time.time() - t0, 2), "oracle_threads": threads,
        "oracle_stage_ms": {s: round(v, 1) for s, v in ms.items()},
    }
    print(name, scale[name]["proof_sha256"], scale[name]["oracle_wall_s"], flush=True)
    with open(path, "w") as fh:
        json.dump(scale, fh, indent=1, sort_keys=True)
