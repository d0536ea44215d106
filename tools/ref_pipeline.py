"""Reference side of the end-to-end check: the compiled reference (oracle/_ref/libpmref.so) runs Propagate::run(0) and Filter::run on
config 1 at half size from the same seeds; prints patch counts, wall time and depth-error quantiles against the ground-truth plane.
Result on the build container (1 core): 1003 seeds -> 27321 patches in 50.9 s; filter removes 0; |z|/scale quantiles 50/90/99 % =
5.98e-4 / 1.29e-3 / 2.16e-3.  (PmMvps::run itself cannot get past iteration 0: Propagate::run exits on its never-drained m_queue.)"""
import sys, time, tempfile, numpy as np
import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mvskit_b200 import synth
from oracle import pyoracle
scene = synth.make_scene(1, scale=0.5).render()
prefix = synth.write_scene(scene, tempfile.mkdtemp(prefix="pm_e2e_"))
ref = pyoracle.RefLib(prefix)
ref.refine_seed(0x5EED0001)
ref.clear_patches(); ref.set_depth(0); ref.create_patches(); ref.set_depth(1)
n0 = ref.collect(0)
t=time.time(); ref.propagate_run(0); t1=time.time()-t
pb = ref.get_patches()
print("seeds", n0, "after propagate", pb.n, "secs", t1)
z = np.abs(pb.coord[:,2])/scene.scene_scale
print("depth err quantiles", np.quantile(z,[0.5,0.9,0.99]), "ncc mean", pb.scal[:,0].mean())
np.savez_compressed('/tmp/ref_after_prop0.npz', coord=pb.coord, normal=pb.normal, scal=pb.scal, images=pb.images, nimages=pb.nimages)
t=time.time(); ref.filter_run(); t2=time.time()-t
pb = ref.get_patches()
z = np.abs(pb.coord[:,2])/scene.scene_scale
print("after filter", pb.n, "secs", t2, "depth err quantiles", np.quantile(z,[0.5,0.9,0.99]))
