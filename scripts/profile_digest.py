"""Turn the outputs of scripts/gpu_profile.sh (gpurun_out/) into the committed summaries under profiles/.
usage: python scripts/profile_digest.py r01c"""
import csv
import io
import json
import os
import subprocess
import sys
from contextlib import redirect_stdout

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "scripts"))
import launch_shares  # noqa: E402
import ncu_summary  # noqa: E402

tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
OUT = os.path.join(ROOT, "gpurun_out")
PROF = os.path.join(ROOT, "profiles")
CMD = "python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu"


def raw_metrics(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    r = list(csv.reader(out.splitlines()))
    return dict(zip(r[0], r[2])), dict(zip(r[0], r[1]))


def capture(fn):
    buf = io.StringIO()
    with redirect_stdout(buf):
        fn()
    return buf.getvalue()


# launch list
csvp = os.path.join(OUT, f"launches_{tag}.csv")
if os.path.exists(csvp):
    txt = capture(lambda: launch_shares.main(csvp, 2))
    plain = json.load(open(os.path.join(OUT, "prof_plain.json")))
    rt = tag.split("_")[0][:3] if tag.startswith("r0") else "r01"
    with open(os.path.join(PROF, f"{rt}_launches.md"), "w") as f:
        f.write(f"# {rt} — ncu launch list of `{CMD}` (the two timed steps)\n\n"
                f"`ncu --metrics gpu__time_duration.sum --clock-control none -k regex:spl_ -s <3 warm-up steps> -c <2 steps>`;\n"
                f"raw csv: `profiles/{rt}_launches.csv`.  Per-launch times under ncu are cold-cache and serialised: compare SHARES.\n"
                f"The same command without ncu printed fit {plain['fit_ms']:.2f} ms + eval {plain['eval_ms']:.2f} ms per step "
                f"(stages: {json.dumps({k: round(v, 3) for k, v in plain['stages_ms'].items()})}).\n\n")
        f.write(txt)
    subprocess.run(["cp", csvp, os.path.join(PROF, f"{rt}_launches.csv")])

traffic = {}
rt = tag.split("_")[0][:3] if tag.startswith("r0") else "r01"
units = {"eval": ("spl_eval_regroup_kernel<3>" if rt != "r01" else "spl_eval_kernel<3>", 1_000_000_000),
         "accumulate": ("spl_moments_kernel", 100_000_000)}
for k in ("eval", "accumulate", "panel", "factor"):
    rep = os.path.join(OUT, f"prof_{tag}_{k}.ncu-rep")
    if not os.path.exists(rep):
        continue
    txt = capture(lambda: ncu_summary.main(rep))
    txt = "\n".join(ln for ln in txt.splitlines()
                    if not any(s in ln for s in ("stalled_drain", "stalled_lg_", "stalled_membar", "stalled_misc",
                                                 "stalled_sleeping", "stalled_tex")))
    with open(os.path.join(PROF, f"{rt}_{k}_bench.md" if rt != "r01" else f"r01_{k}.md"), "w") as f:
        kn = "spl_moments_kernel (the accumulate stage)" if k == "accumulate" else f"spl_{k}*"
        f.write(f"# {rt} — `ncu --set full --clock-control none` of {kn} inside `{CMD}` (bench size)\n\n```\n{txt}\n```\n")
    if k in units:
        m, _ = raw_metrics(rep)
        rd = float(m["dram__bytes_read.sum"].replace(",", ""))
        wr = float(m["dram__bytes_write.sum"].replace(",", ""))
        # ncu prints these in the unit of the second header row; re-read with base units
        out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv", "--print-units", "base"], capture_output=True,
                             text=True).stdout
        r = list(csv.reader(out.splitlines()))
        mm = dict(zip(r[0], r[2]))
        rd = float(mm["dram__bytes_read.sum"].replace(",", ""))
        wr = float(mm["dram__bytes_write.sum"].replace(",", ""))
        traffic[units[k][0]] = {"dram_bytes": rd + wr, "dram_bytes_read": rd, "dram_bytes_write": wr, "units": units[k][1],
                                "source": f"ncu --set full of `{CMD}`, one launch"}
if traffic:
    json.dump(traffic, open(os.path.join(PROF, f"{rt}_traffic.json"), "w"), indent=1)
    print(json.dumps(traffic, indent=1))
