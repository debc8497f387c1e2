import json,sys
d=json.loads(sys.stdin.read()); print("%s value %.0f p50 %.4f ms/step %.4f" % (sys.argv[1], d["value"], d["p50_ms"], d["ms_per_step"]))
