#!/usr/bin/env python3
"""Extracts the per-modulus constants that the reference vendors in its seven
generated `zp` packages (q, qInvNeg, rSquare, Limbs, Bits) into
tests/golden/field_constants.json.  Run in the build container only
(/root/reference is not present on the GPU box); the JSON is committed.

Sources: <pkg>/element.go  `q[base16]` comment, `const qInvNeg`, `var rSquare`.
"""
import json, re, pathlib

REF = pathlib.Path("/root/reference")
PKGS = {
    "zp110": "buckler/internal/zp110",
    "zp128": "examples/mult/zp",
    "zp220": "buckler/internal/zp220",
    "zp240": "examples/bfv/zp",
    "zp255": "jindo/internal/zp",
    "zp440": "buckler/internal/zp440",
    "zp880": "buckler/internal/zp880",
}

out = {}
for name, rel in PKGS.items():
    src = (REF / rel / "element.go").read_text()
    q = int(re.search(r"q\[base16\] = 0x([0-9a-f]+)", src).group(1), 16)
    limbs = int(re.search(r"Limbs = (\d+)", src).group(1))
    bits = int(re.search(r"Bits  = (\d+)", src).group(1))
    qinvneg = int(re.search(r"const qInvNeg = (\d+)", src).group(1))
    rs = re.search(r"var rSquare = Uint\{([^}]*)\}", src).group(1)
    rsq = [int(x) for x in re.findall(r"\d+", rs)]
    qlimbs = [int(m) for m in re.findall(r"^\s+q\d+\s+= (\d+)", src, re.M)][:limbs]
    assert sum(v << (64 * i) for i, v in enumerate(qlimbs)) == q
    out[name] = {
        "source": f"{rel}/element.go",
        "q_hex": hex(q), "limbs": limbs, "bits": bits,
        "qInvNeg": str(qinvneg), "rSquare": [str(x) for x in rsq],
    }
path = pathlib.Path(__file__).with_name("field_constants.json")
path.write_text(json.dumps(out, indent=1) + "\n")
print("wrote", path, list(out))
