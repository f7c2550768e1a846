set -x
mkdir -p gpurun_out
timeout 800 python -m pytest tests -m gpu -q 2>&1 | tail -8
# output-step cost at 8 M elements: full integration-point download vs the device-side column split
timeout 600 python - <<'PY' > gpurun_out/u_output_step.txt 2>&1
import time, numpy as np
from nimblesm_b200 import capi
from nimblesm_b200.mesh import structured_cube
n=200
mesh=structured_cube(n)
c=capi.Context(0); c.set_nodes(mesh["x"],mesh["y"],mesh["z"]); c.add_block(1,mesh["conn"][1],"neohookean",1.6e12,0.8e12,7.8); c.finalize(capi.ASSEMBLY_ATOMIC,2)
c.compute_lumped_mass()
v=np.zeros((len(mesh["x"]),3)); v[:,0]=1000.0*mesh["x"]; c.upload("velocity",v)
t=c.step(3,0.0,2.2e-9,store_ipt_last=True)
for rep in range(2):
    t0=time.perf_counter(); full=c.element_data(1); t1=time.perf_counter()
    six=c.element_components(1,[9,10,11,12,13,14]); t2=time.perf_counter()
    der=c.derived_element_data(1); t3=time.perf_counter()
    print("8M elements: full [n][8][15] download %.3f s (%.2f GB); 6 components split on the device %.3f s (%.2f GB); derived (volume + 15 averages) %.3f s"%(t1-t0, full.nbytes/1e9, t2-t1, six.nbytes/1e9, t3-t2))
assert np.array_equal(six[0], full.reshape(-1,120)[:,9])
PY
cat gpurun_out/u_output_step.txt
