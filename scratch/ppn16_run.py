"""bf16 PPN entry point: parity smoke + microbench (run on the GPU box)."""
import sys, json, torch
sys.path.insert(0, '.')
import bench
print(json.dumps(bench.ppn_microbench("cuda", bench.peaks()), indent=1))
