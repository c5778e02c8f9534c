"""Generates tests/golden/e2e_proofs.json from the pure-Python restatement of the reference
(oracle/pyref.py): the two reference e2e configurations (tests/e2e_goldilocks.rs, tests/e2e_babybear.rs).
The reference itself (Rust) cannot run in this image, so these are ORACLE outputs, not reference
outputs ("parity unpinned" for the transcript-dependent bytes)."""
import hashlib
import json
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", ".."))
from oracle import pyref as R  # noqa: E402

out = {}
for F, steps in ((R.Goldilocks, 9), (R.BabyBear, 7)):
    claim = R.FibonacciClaim(F, steps)
    trace = claim.trace(2)
    cfg = R.StarkConfig(F, 20, 2, trace.step_number(), trace.constrain_number())
    proof = R.Stark(cfg).prove(claim, 2)
    raw = R.serialize_proof(F, proof)
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
with open(os.path.join(os.path.dirname(__file__), "e2e_proofs.json"), "w") as fh:
    json.dump(out, fh, indent=1, sort_keys=True)
print(json.dumps(out, indent=1, sort_keys=True))
