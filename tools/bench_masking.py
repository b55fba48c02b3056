"""Masking microbench only (BASELINE.json configs[3]); prints the table bench.py embeds in its JSON line."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
import cooperative_training_and_latent_space_data_augmentation_b200 as pkg  # noqa: E402

roofline, sweep = bench.masking_microbench(pkg, bench.measured_peaks(), iters=int(sys.argv[1]) if len(sys.argv) > 1 else 30)
print(json.dumps({"roofline": roofline, "sweep": {k: {"us": round(v["us"], 2), "GBps": round(v["GBps"], 1)} for k, v in sweep.items()}}))
