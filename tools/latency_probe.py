"""Config-1 latency (one ~2 s utterance, batch 1, 10 NFE, synthesise + HiFT, host in / host out) under switch variants.
Each variant runs in its own process (the switches are read once).  usage: python tools/latency_probe.py [tokens]"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
VARIANTS = [{}, {"JYUTVOICE_B200_MLP": "0"}, {"JYUTVOICE_B200_PAIR": "0"}, {"JYUTVOICE_B200_MLP": "0", "JYUTVOICE_B200_PAIR": "0"},
            {"JYUTVOICE_B200_ATTN_NCH": "3"}, {"JYUTVOICE_B200_GRAPH": "0"}, {"JYUTVOICE_B200_WRES": "0"}, {"JYUTVOICE_B200_PDL": "0"},
            {"JYUTVOICE_B200_EPI16": "0"}, {"JYUTVOICE_B200_MLP": "0", "JYUTVOICE_B200_EPI16": "0"}]
if os.environ.get("LATENCY_PROBE_VARIANTS"):  # e.g. "0,1,8,9"
    VARIANTS = [VARIANTS[int(i)] for i in os.environ["LATENCY_PROBE_VARIANTS"].split(",")]

if len(sys.argv) > 1 and sys.argv[1] == "--child":
    sys.path.insert(0, ROOT)
    import torch
    import bench
    tokens = int(sys.argv[2])
    sds = bench.all_state_dicts()
    dev = torch.device("cuda", 0)
    print(json.dumps({"p50_ms": bench.small_latency("bf16", sds, dev, tokens=tokens, reps=15)}))
else:
    tokens = sys.argv[1] if len(sys.argv) > 1 else "16"
    for v in VARIANTS:
        env = dict(os.environ, **v)
        out = subprocess.run([sys.executable, __file__, "--child", tokens], env=env, capture_output=True, text=True)
        line = [l for l in out.stdout.splitlines() if l.startswith("{")]
        print(v or "default", line[-1] if line else out.stderr[-400:])
