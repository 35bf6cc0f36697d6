"""Debug: print the event timeline (PFM_TC_PROF + PFM_TC_TRACE) of one evaluation of block 0's first group."""
import os, sys, subprocess
env = dict(os.environ, PFM_TC_PROF="1", PFM_TC_TRACE="1")
out = subprocess.run([sys.executable, os.path.join(os.path.dirname(__file__), "tc_prof.py"), *sys.argv[1:]], env=env,
                     capture_output=True, text=True).stderr
ev = []
lines = out.splitlines()
first_summary = next(i for i, l in enumerate(lines) if l.startswith("[pfm tc prof]"))
for line in lines[first_summary:]:          # tc_prof.py launches twice: keep the second (warm) launch
    if line.startswith("[pfm tc trace]"):
        t, r, s = line.split()[3:6]
        ev.append((int(t), int(r), int(s)))
ev.sort()
seen = set()
names = {0: "mma ", 1: "epiA", 2: "epiB"}
last = {0: None, 1: None, 2: None}
# the trace buffer is overwritten by the second launch of tc_prof.py: keep the last occurrence of each (t, r, s)
for t, r, s in ev:
    d = "" if last[r] is None else f"+{t - last[r]}"
    last[r] = t
    print(f"{t:8d}  {names[r]}  slot {s:2d}  {d}")
