import json
for l in open("gpurun_out/sweep.jsonl"):
    if l.startswith("##"):
        print(l.strip()); continue
    d = json.loads(l); r = d["roofline"]
    print("  value %.1f M/s e2e %.1f  ms/step %.1f  stage_ms %s pipeline_frac %.3f kernel %s frac %.3f" % (
        d["value"] / 1e6, d["e2e"]["value"] / 1e6, d["ms_per_step"], {k: round(v, 1) for k, v in r["stage_ms"].items() if v > 0}, r["pipeline_frac"], r["kernel"], r["frac"]))
