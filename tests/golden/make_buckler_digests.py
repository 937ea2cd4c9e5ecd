"""Writes tests/golden/buckler_proof_digests.json: SHA-256 of the restated Buckler prover's proofs (oracle/buckler_ref.py) for the four
test circuits at N = 256, on the seeded witness / CRS / randomness tape that tests/test_oracle_buckler.py::test_prove_verify uses.

What this pins: the oracle against itself over time (a regression in the restatement, the samplers or the transcripts changes the
digest) and, through tests/test_gpu_buckler.py, the GPU provers against the same bytes.  What it does NOT pin: the reference -- its
proofs depend on lattigo / gnark-crypto conventions that cannot be checked offline ("unpinned").

    python tests/golden/make_buckler_digests.py"""
import hashlib
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))


def proof_digest(pf, L):
    """canonical bytes: every commitment / proof polynomial coefficient as u64 LE in order, field elements as 8 L bytes big-endian"""
    h = hashlib.sha256()

    def polys(x):
        if isinstance(x, int):
            h.update(int(x).to_bytes(8, "little"))
        else:
            for y in x:
                polys(y)
    polys(pf["Witness"])
    for e in pf["Evals"]:
        h.update(int(e).to_bytes(8 * L, "big"))
    for k in ("LinCheckMaskSum", "SumCheckMaskSum"):
        h.update(b"\x00" if pf.get(k) is None else b"\x01" + int(pf[k]).to_bytes(8 * L, "big"))
    for k in ("Partial", "PartialMask", "Encode", "MLWE", "InCommit"):
        polys(pf["EvalProof"][k])
    return h.hexdigest()


def main():
    from oracle import buckler_ref as B, jindo_ref as J, pyref
    from test_oracle_buckler import circuit_and_witness
    out = {}
    for circuit, name in (("pk", "zp220"), ("mult", "zp128"), ("ct", "zp240"), ("norm", "zp220")):
        q = pyref.modulus(name)
        (define, n_pw, n_w), (pw, w) = circuit_and_witness(circuit, q, 256, 3)
        tape = J.Tape(b"buckler-" + circuit.encode())
        C = B.compile_circuit(q, 256, define, n_pw, n_w, b"Buckler!", tape)
        pf = B.Prover(C, tape).prove(pw, w)
        out[circuit] = {"field": name, "rank": 256, "witness_seed": 3, "tape_seed": "buckler-" + circuit, "crs": "Buckler!",
                        "commitments": len(pf["Witness"]), "sha256": proof_digest(pf, (q.bit_length() + 63) // 64)}
        print(circuit, out[circuit]["sha256"])
    with open(os.path.join(HERE, "buckler_proof_digests.json"), "w") as fh:
        json.dump(out, fh, indent=1)


if __name__ == "__main__":
    main()
