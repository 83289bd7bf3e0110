"""Generates tests/golden/join_split_proof.json from the UNMODIFIED reference prover (oracle/_ref/js_prover_cpu, built by
`make -C oracle prover` from /root/reference; see oracle/js_harness.cpp).  Run in the dev container:

    python tests/golden/gen_prover_golden.py

The fixture pins the end-to-end parity case of BASELINE.json config #4: the 1 952-byte TurboPLONK join-split proofs that
the reference CPU prover emits for the first two deterministic noop transactions.  tests/test_gpu_prover.py demands the
same bytes from the binary whose MSM / FFT entry points are resolved to the CUDA library.
"""
import hashlib
import json
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.path.join(ROOT, "oracle", "_ref")


def main():
    out = subprocess.run([os.path.join(REF, "js_prover_cpu"), os.path.join(REF, "srs_db"), "2"], check=True,
                         capture_output=True, text=True).stdout.strip().splitlines()[-1]
    r = json.loads(out)
    assert r["verified"] and r["proof_bytes"] == 1952
    fx = {"n": r["n"], "proof_bytes": r["proof_bytes"], "first_proof": r["first_proof"], "last_proof": r["last_proof"],
          "first_proof_sha256": hashlib.sha256(bytes.fromhex(r["first_proof"])).hexdigest(),
          "generator": "oracle/_ref/js_prover_cpu srs_db 2 (reference CPU prover, deterministic get_engine)"}
    with open(os.path.join(HERE, "join_split_proof.json"), "w") as f:
        json.dump(fx, f, indent=1)
    print("wrote join_split_proof.json", fx["first_proof_sha256"])


if __name__ == "__main__":
    main()
