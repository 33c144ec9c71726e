"""Secondary measurement (BASELINE.json configs[2]): examples/field_only -- vacuum Maxwell on 256^3 with the soft
plane-wave source and the MABC boundary in x, schedule of examples/field_only/main.cpp:142-145
(Theta_E(dt/2), source, Theta_B(dt), Theta_E(dt/2)).  Prints one JSON line: cell-updates/s and the HBM roofline of
the curl sweeps (72 B per cell and sweep, 3 sweeps per step: SURVEY 8d)."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import strugepic_b200 as spic  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 200
sim = spic.Simulation((n, n, n), periodic=(0, 1, 1), interp=spic.P8R2)
sim.set_uniform_field(spic.FIELD_E, [0, 0, 0])
sim.set_uniform_field(spic.FIELD_B, [0, 0, 0])
stream = torch.cuda.ExternalStream(sim.stream())
dt = 0.5
for s in range(20):
    sim.field_only_step(4, 1, 0.1, 0.02, dt, s)
sim.sync()
sim.set_option("time_kernels", 1)
sim.kernel_times(reset=True)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(stream)
for s in range(20, 20 + steps):
    sim.field_only_step(4, 1, 0.1, 0.02, dt, s)
e1.record(stream)
sim.sync()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1)
kt = sim.kernel_times(reset=True)
cells = n ** 3
peak = 6650.0
p = os.path.join(ROOT, "MEASURED_PEAKS.json")
src = "fallback (B200_PROFILING.md)"
if os.path.isfile(p):
    peak, src = json.load(open(p))["hbm_gbs"], "measured (MEASURED_PEAKS.json)"
curl_ms, curl_n = kt["curl"]
ach = 72.0 * cells / (curl_ms / max(curl_n, 1) * 1e-3) / 1e9
print(json.dumps({
    "metric": "cell-updates/s (field_only: source + MABC x, 3 curl sweeps per step)", "value": cells * steps / (ms * 1e-3),
    "unit": "cell-updates/s", "n_gpus": 1, "steps": steps, "ms_per_step": ms / steps, "dtype": "f64",
    "config": {"workload": "examples/field_only %d^3, x non-periodic, source plane i=4 comp Y E0=0.1 omega=0.02, MABC x" % n},
    "roofline": {"kernel": "k_curl", "bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                 "peak_source": src, "avg_launch_ms": curl_ms / max(curl_n, 1), "launches_timed": curl_n,
                 "algorithmic_bytes_per_launch": 72.0 * cells},
    "whole_step_gbs": 216.0 * cells * steps / (ms * 1e-3) / 1e9,
    "field_energy_after": sim.get_total_energy()[0]}))
