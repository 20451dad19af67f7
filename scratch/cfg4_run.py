"""config-4 end-to-end leg alone (run on the GPU box)."""
import sys, json, torch
sys.path.insert(0, '.')
import bench
dev = torch.device("cuda", 0)
flush = torch.empty(64 * 1024 * 1024, dtype=torch.float32, device=dev)
print(json.dumps(bench.config4_e2e_bench(dev, flush, torch.cuda.current_stream()), indent=1))
