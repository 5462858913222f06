import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch, util, ref_ext
from oracle import splat_oracle
from s3gaussian_b200 import synthetic as syn
from s3gaussian_b200 import diff_gaussian_rasterization as ours
cloud, cam = syn.make_small_scene(P=300, width=96, height=64, seed=21)
cloud.opacity[:60] = -8.0; cloud.opacity[60:120] = 12.0
cloud.scaling[120:180, 0] += 3.0; cloud.scaling[120:180, 1] -= 2.0
d = util.scene_inputs(cloud, cam, mode="rgb")
gc, gd = util.seeded_grads(d, 5)
o = util.oracle_run(splat_oracle, d, gc, gd)
ref = ref_ext.load()
r = util.run_module(ref, d, "cuda:0", gc, gd)
for trial in range(2):
    m = util.run_module(ours, d, "cuda:0", gc, gd)
    for k, ok in (("opacities","opacity"),("means3D","means3D"),("scales","scales"),("rotations","rotations"),("colors_precomp","colors")):
        a = m["grads"][k].cpu().numpy(); b = r["grads"][k].cpu().numpy(); c = o["grads"][ok].reshape(a.shape)
        print(k, "ours-ref %.2e  ours-oracle %.2e  ref-oracle %.2e" % (util.relerr(a,b), util.relerr(a,c), util.relerr(b,c)))
r2 = util.run_module(ref, d, "cuda:0", gc, gd)
print("ref run-to-run opacities", util.relerr(r2["grads"]["opacities"].cpu().numpy(), r["grads"]["opacities"].cpu().numpy()))
