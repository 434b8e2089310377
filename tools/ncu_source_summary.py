"""Aggregates the per-instruction page of an ncu report (`ncu -i rep --page source --csv`) by opcode: share of executed
instructions, share of stall samples and the leading stall reasons.  usage: python tools/ncu_source_summary.py source.csv out.json"""
import collections
import csv
import json
import re
import sys

STALLS = ("stall_wait", "stall_math", "stall_long_sb", "stall_dispatch", "stall_not_selected", "stall_selected", "stall_barrier",
          "stall_short_sb", "stall_mio", "stall_lg", "stall_no_inst")


def main(src, out):
    rows = list(csv.reader(open(src)))
    h = next(i for i, r in enumerate(rows) if "Source" in r and any("Samples" in c for c in r))
    ix = {name: i for i, name in enumerate(rows[h])}
    agg = collections.defaultdict(collections.Counter)
    for r in rows[h + 1:]:
        if len(r) < len(rows[h]):
            continue
        m = re.match(r"(@!?U?P\d+\s+)?([A-Z0-9_.]+)", r[ix["Source"]].strip())
        if not m:
            continue
        op = m.group(2)
        key = "IMAD.WIDE" if op.startswith("IMAD.WIDE") else op.split(".")[0]
        try:
            agg[key]["executed"] += int(r[ix["Instructions Executed"]] or 0)
            agg[key]["samples"] += int(r[ix["# Samples"]] or 0)
            for st in STALLS:
                agg[key][st] += int(r[ix[st]] or 0)
        except (ValueError, KeyError):
            continue
    te = sum(v["executed"] for v in agg.values()) or 1
    ts = sum(v["samples"] for v in agg.values()) or 1
    ops = []
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1]["executed"]):
        if v["executed"] * 200 < te:
            continue
        ops.append({"opcode": k, "executed_pct": round(100.0 * v["executed"] / te, 2), "stall_samples_pct": round(100.0 * v["samples"] / ts, 2),
                    "stall_reasons_pct": {s[6:]: round(100.0 * v[s] / max(1, v["samples"]), 1) for s in STALLS if v[s] * 20 >= max(1, v["samples"])}})
    json.dump({"source": src, "warp_instructions_executed": te, "stall_samples": ts, "by_opcode": ops}, open(out, "w"), indent=1)
    print("wrote", out)


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
