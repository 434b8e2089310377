#!/usr/bin/env python3
"""Extracts the known-answer vectors the reference's own tests hold for the MSM/NTT path into
tests/golden/reference_kats.json.  Run in the build container (reads /root/reference); the JSON
is committed because /root/reference does not exist on the GPU box.

Sources (relative to /root/reference):
  aiken-verifier/aiken_halo2/lib/transcript.ak:108-382   encodings, transcript challenges, golden proof
  aiken-verifier/aiken_halo2/lib/omega_rotations.ak:48-81 omega(k=14), omega^-1, rotations
  aiken-verifier/aiken_halo2/lib/lagrange.ak:133-187      Lagrange basis values, barycentric weight 1/n
  plinth-verifier/plutus-halo2/test/ProofData.hs          H2MO fixture (points, evaluations, challenges)
  plinth-verifier/plutus-halo2/test/Halo2MultiOpenMSM.hs:26-42  expected v, f_eval, q_eval_sets
  plinth-verifier/plutus-halo2/src/Plutus/Crypto/BlsTypes.hs:97,102-103, Constants.hs:10-13, Transcript.hs:78-79
"""
import json
import re
import sys

REF = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"


def read(p):
    return open(REF + "/" + p).read()


out = {}

# ---------------------------------------------------------------- ProofData.hs
pd = read("plinth-verifier/plutus-halo2/test/ProofData.hs")
scalars = {m.group(1): int(m.group(2), 16) for m in re.finditer(r"^(\w+) = mkScalar (0x[0-9a-fA-F]+)", pd, re.M)}
points = {}
for m in re.finditer(r"^(\w+) =\s*\n\s*\(constructG1Point \. bimap mkFp mkFp\)\s*\n\s*\(\s*(0x[0-9a-fA-F]+)\s*\n\s*,\s*(0x[0-9a-fA-F]+)",
                     pd, re.M):
    points[m.group(1)] = [int(m.group(2), 16), int(m.group(3), 16)]
cm_src = pd[pd.index("commitmentMap =\n"):pd.index("-- evaluation of parts from buildMSM")]
cmap = []
for m in re.finditer(r"\((\w+), (\d+), \[([\w, ]+)\], \[([\w, ]+)\]\)", cm_src):
    cmap.append({"commitment": m.group(1), "set": int(m.group(2)),
                 "points": [s.strip() for s in m.group(3).split(",")],
                 "evals": [s.strip() for s in m.group(4).split(",")]})
assert len(cmap) == 12, len(cmap)
out["h2mo"] = {
    "source": "plinth-verifier/plutus-halo2/test/ProofData.hs; expected values Halo2MultiOpenMSM.hs:26-42",
    "scalars": {k: hex(v) for k, v in scalars.items()},
    "points": {k: [hex(v[0]), hex(v[1])] for k, v in points.items()},
    "commitment_map": cmap,
    "point_sets": [["x_current", "x_next"], ["x_current"], ["x_current", "x_next", "x_last"]],
    "proof_x3_q_evals": ["q_eval_on_x3_1", "q_eval_on_x3_2", "q_eval_on_x3_3"],
}
ex = read("plinth-verifier/plutus-halo2/test/Halo2MultiOpenMSM.hs")
out["h2mo"]["expected_v"] = re.search(r"expectedV = mkScalar (0x[0-9a-f]+)", ex).group(1)
out["h2mo"]["expected_f_eval"] = re.search(r"expectedFEval = mkScalar (0x[0-9a-f]+)", ex).group(1)
sets_src = ex[ex.index("expectedQEvalSets =\n"):ex.index("assertCorrectV ::")]
qs = []
for grp in re.split(r"\n    ,", sets_src.split("=", 1)[1]):
    vals = re.findall(r"mkScalar (0x[0-9a-f]+)", grp)
    if vals:
        qs.append(vals)
assert [len(x) for x in qs] == [2, 1, 3], qs
out["h2mo"]["expected_q_eval_sets"] = qs

# ---------------------------------------------------------------- transcript.ak
tr = read("aiken-verifier/aiken_halo2/lib/transcript.ak")
proof_hex = re.search(r'let proof_for_testing =\s*\n\s*#"([0-9a-f]{2240})"', tr).group(1)
full = tr[tr.index("test full_proof_deserialization_for_simple_mul_circuit"):]
chal = {}
for name in ("gamma", "y", "x", "adviceEval1", "adviceEval2", "adviceEval3", "x1", "x2", "x3", "x4"):
    m = re.search(r"expect\s*\n?\s*%s == from_int\(\s*\n?\s*(0x[0-9a-f]+)" % name, full)
    chal[name] = m.group(1)
out["transcript"] = {
    "source": "aiken-verifier/aiken_halo2/lib/transcript.ak:108-382",
    "R256": re.search(r"(0x1824b159[0-9a-f]+)", tr).group(1),
    "repr_only": {"repr": "0x53772fda8c4d27d16e6d1b3b0ed0f0c492414695f8050480aaeb9f0c1257bc6b",
                  "challenge": re.search(r"squeeze == from_int\(\s*\n\s*(0x[0-9a-f]+)", tr).group(1)},
    "after_scalar_42": re.search(r"test adding_scalar_to_transcript.*?challenge == from_int\(\s*\n\s*(0x[0-9a-f]+)", tr, re.S).group(1),
    "after_point_42G": re.search(r"test adding_g1_to_transcript.*?challenge == from_int\(\s*\n\s*(0x[0-9a-f]+)", tr, re.S).group(1),
    "mixed": {
        "proof": re.search(r"test squeeze_challenge_calculations.*?#\"([0-9a-f]+)\"", tr, re.S).group(1),
        "scalar": "0x71eda753299d7d483339d80809a1d80553bda442fffe5bfeffffffff00000001",
        "challenge": re.search(r"test squeeze_challenge_calculations.*?challenge == from_int\(\s*\n\s*(0x[0-9a-f]+)", tr, re.S).group(1),
    },
    "golden_proof": {"repr": "0x53772fda8c4d27d16e6d1b3b0ed0f0c492414695f8050480aaeb9f0c1257bc6b",
                     "public_inputs": [42, 42, 42], "proof": proof_hex, "expected": chal,
                     "pi_compressed": re.search(r'decompress\(\s*\n\s*#"([0-9a-f]{96})"', full).group(1)},
}
out["g1_encoding"] = {
    "source": "aiken-verifier/aiken_halo2/lib/transcript.ak:121-179",
    "generator": re.search(r'let generator_bytes =\s*\n\s*#"([0-9a-f]{96})"', tr).group(1),
    "neg_generator": re.search(r'let negated_generator_bytes =\s*\n\s*#"([0-9a-f]{96})"', tr).group(1),
    "g_times_42": re.search(r'let point_bytes =\s*\n\s*#"([0-9a-f]{96})"', tr).group(1),
    "scalar_r_bytes": "01000000fffffffffe5bfeff02a4bd5305d8a10908d83933487d9d2953a7ed73",
    "scalar_overflow_bytes": "01000000fffffffffe5bfeff42a4bd5305d8a10908d83933487d9d2953a7ed71",
    "scalar_overflow_value": "0x71eda753299d7d483339d80809a1d80553bda442fffe5bfeffffffff00000001",
}

# ---------------------------------------------------------------- omega / lagrange
om = read("aiken-verifier/aiken_halo2/lib/omega_rotations.ak")
t = om[om.index("test calculate_rotations"):]
vals = re.findall(r"from_int\(\s*\n?\s*(0x[0-9a-f]+|1)\s*,?\s*\n?\s*\)", t)
out["omega_k14"] = {"source": "aiken-verifier/aiken_halo2/lib/omega_rotations.ak:48-81",
                    "omega": vals[0], "omega_inv": vals[1], "rotations_m6_to_0": vals[2:9]}
lg = read("aiken-verifier/aiken_halo2/lib/lagrange.ak")
t = lg[lg.index("test calculate_lagrange_basis"):]
vals = re.findall(r"from_int\(\s*\n?\s*(0x[0-9a-f]+|1)\s*,?\s*\n?\s*\)", t)
out["lagrange_k14"] = {"source": "aiken-verifier/aiken_halo2/lib/lagrange.ak:133-187",
                       "x": vals[0], "xn": vals[1], "barycentric_weight": vals[2], "rotations": vals[3:10],
                       "expected": vals[10:17]}
out["constants"] = {
    "source": "BlsTypes.hs:97,102-103; Constants.hs:10-13",
    "r": hex(int(re.search(r"bls12_381_field_prime = (\d+)", read("plinth-verifier/plutus-halo2/src/Plutus/Crypto/BlsTypes.hs")).group(1))),
    "p": "0x" + re.search(r"0x(1a0111ea[0-9a-f]+)", read("aiken-verifier/aiken_halo2/lib/bls_utils.ak")).group(1),
    "delta": re.search(r"(0x0?8634d0aa[0-9a-f]+)", read("plinth-verifier/plutus-halo2/src/Plutus/Crypto/Constants.hs"), re.I).group(1),
}
json.dump(out, open(__file__.rsplit("/", 1)[0] + "/reference_kats.json", "w"), indent=1)
print("wrote reference_kats.json:", {k: len(v) for k, v in out.items()})
