import json,sys
for f in sys.argv[1:]:
    try:
        d=json.load(open(f)); print(f.split('/')[-1], "value %.0f e2e %.0f conv %.3f ss %.3f sustained %s" % (d["value"], d["e2e"]["value"], d["roofline"]["conv_ms_per_step"], d["single_stream"]["ms_per_step"], d.get("sustained",{}).get("value")))
    except Exception as e: print(f, "ERR", e)
