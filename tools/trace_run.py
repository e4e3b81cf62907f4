import sys; sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import numpy as np, multih_b200 as m
sc = m.scenes.make_scene(3000, 6, seed=7)
ctx = m.Context(m.capi.default_params(locality=1/20.0))
lab,H,K = ctx.process(sc.pts, sc.aff, sc.F)
print("K",K,"it",ctx.iterations)
