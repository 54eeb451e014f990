"""Headline numbers of a bench.py JSON line: python scripts/show_bench.py file"""
import json, sys
d = json.loads([l for l in open(sys.argv[1]) if l.startswith("{")][-1])
print("value %.2f M/s (%.4f ms/step)  roofline %.3f (%s %.4f ms)  e2e %.2f M/s  two-in-flight %.2f  commands %.2f  cpu %.0f/s  launches %d  clocks %s" % (
    d["value"] / 1e6, d["ms_per_step"], d["roofline"]["frac"], d["roofline"]["kernel"], d["roofline"]["kernel_ms"], d["e2e"]["value"] / 1e6,
    d["e2e"]["two_in_flight"]["value"] / 1e6, d["e2e"]["commands"]["value"] / 1e6, d["cpu_baseline"]["value"], d["gpu_launches"], d["clocks"]))
