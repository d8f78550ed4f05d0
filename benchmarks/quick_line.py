"""dev helper: one-line summary of a bench.py JSON line read from stdin (value M samples/s, roofline frac, forward M samples/s)"""
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print(sys.argv[1], round(d["value"]/1e6), round(d["roofline"]["frac"],4), round(d["forward_only"]["value"]/1e6))
