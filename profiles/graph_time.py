import sys, time
sys.path.insert(0, "/root/repo")
import numpy as np
import __graft_entry__ as g
pkg = g.load_package()
"""Launch-bound case (BASELINE config 1): 1000 RK4 steps on the 1 542-triangle square, eager launches vs the CUDA graph of a step."""
import gzip, tempfile
from pathlib import Path
tmp = Path(tempfile.mkdtemp()) / "square.msh"
tmp.write_bytes(gzip.decompress(Path("/root/repo/tests/golden/meshes/square.msh.gz").read_bytes()))
model = pkg.Model.open_msh(tmp, 1)
cfg = model.parse_config(Path("/root/repo/tests/golden/configs/square_pulse.conf"))
mesh = pkg.Mesh(model, cfg)
u0 = mesh.initial_condition()
kernels = [int(a) for a in sys.argv[1:]] or [0]
for kernel, mode in [(k, m) for k in kernels for m in (0, 1, 0, 1)]:
    eng = pkg.Engine(mesh, options={"kernel": kernel})
    eng.set_option("graph", mode)
    eng.set_state(u0)
    eng.run(pkg.RUNGE_KUTTA, 0.0, 50)
    t0 = time.perf_counter()
    eng.run(pkg.RUNGE_KUTTA, 0.0, 1000)
    dt = time.perf_counter() - t0
    print(f"config 1 (square.msh p=1, K={mesh.K}), 1000 RK4 steps, graph={mode}: host {dt*1e3:.2f} ms, device {eng.last_run_ms:.2f} ms, kernel {eng.kernel_name}")
    eng.close()
