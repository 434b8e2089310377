"""The headline numbers quoted in README.md are the ones in the committed bench lines under profiles/ (no GPU needed)."""
import json
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _line(name):
    return json.loads(open(os.path.join(ROOT, "profiles", name)).read().strip().splitlines()[-1])


def test_readme_headline_matches_committed_bench_lines():
    readme = open(os.path.join(ROOT, "README.md")).read()
    n1 = _line("r02_bench_n1_default.json")
    assert n1["metric"] == "g1_msm_points_per_s" and n1["n_gpus"] == 1 and n1["parity"]["parity_full"] is True
    assert "%.1f ms" % n1["ms_per_step"] in readme, n1["ms_per_step"]
    assert "%.1f ms" % n1["e2e"]["ms_per_step"] in readme
    assert "%.3f ms" % n1["ntt"]["ms_per_step"] in readme
    assert 0.85 < n1["roofline"]["frac"] < 0.95 and "90 %" in readme
    for n, fmt in ((2, "%.1f ms"), (4, "%.1f ms"), (8, "%.2f ms")):
        d = _line("r02_bench_n%d.json" % n)
        assert d["n_gpus"] == n and d["parity"]["dlog_identity"] is True
        assert fmt % d["ms_per_step"] in readme, (n, d["ms_per_step"])
        sp = d["single_process"]
        assert sp["parity_with_per_rank_path"] is True and sp["batched_verify"]["accepted_s_left_eq_right"] in (True, None)
        assert "%.1f ms" % sp["e2e_single_process"]["ms_per_step"] in readme, n
    # the proof traces of the 1-GPU line: every commitment was compared with the checker
    for t in n1["circuits"]["traces"]:
        assert t["parity"] is True and t["parity_commitments_checked"] == t["commitments"]
    assert n1["batched_verify"]["accepted_s_left_eq_right"] is True
