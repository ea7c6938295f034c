"""Measure the experimental switches one at a time (round-2 opener): runs tools/time_breakdown.py in a fresh process
per setting and prints the steady-state step time next to the baseline.

  python tools/flag_sweep.py [batch]        (one B200, ~25 s per setting)
"""
import os
import re
import subprocess
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SETTINGS = [
    ("baseline", {}),
    ("CRIS_B200_BN_STREAM=0 (register-streaming BatchNorm kernels)", {"CRIS_B200_BN_STREAM": "0"}),
    ("CRIS_B200_BIAS_MMA=1", {"CRIS_B200_BIAS_MMA": "1"}),
    ("CRIS_B200_HALO_CONV=1", {"CRIS_B200_HALO_CONV": "1"}),
    ("CRIS_B200_HALO_CONV=1 CRIS_B200_BIAS_MMA=1", {"CRIS_B200_HALO_CONV": "1", "CRIS_B200_BIAS_MMA": "1"}),
]
if os.environ.get("CRIS_SWEEP"):  # comma-separated subset / custom settings: NAME=VAL+NAME=VAL,...
    SETTINGS = [("baseline", {})] + [(x, dict(kv.split("=") for kv in x.split("+"))) for x in os.environ["CRIS_SWEEP"].split(",")]


def main():
    batch = sys.argv[1] if len(sys.argv) > 1 else "64"
    rows = []
    for name, env in SETTINGS:
        e = dict(os.environ)
        e.update(env)
        r = subprocess.run([sys.executable, os.path.join(REPO, "tools", "time_breakdown.py"), batch], env=e,
                           capture_output=True, text=True, timeout=600)
        last = [l for l in r.stdout.splitlines() if l.startswith("it ")]
        if r.returncode != 0 or not last:
            rows.append((name, None, (r.stderr or r.stdout)[-300:]))
            continue
        m = re.search(r"gpu fwd\s+([\d.]+) bwd\s+([\d.]+) opt\s+([\d.]+).*total\s+([\d.]+) ms", last[-1])
        rows.append((name, tuple(float(x) for x in m.groups()), ""))
    base = next((v for n, v, _ in rows if n == "baseline" and v), None)
    for name, v, err in rows:
        if v is None:
            print(f"{name:45s} FAILED {err}")
        else:
            d = f"{v[3] - base[3]:+.2f} ms" if base else ""
            print(f"{name:45s} fwd {v[0]:6.2f} bwd {v[1]:6.2f} opt {v[2]:5.2f} total {v[3]:6.2f} ms  {d}")


if __name__ == "__main__":
    main()
