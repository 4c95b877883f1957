import sys, os, numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(__file__), '..', 'tests')); sys.path.insert(0, os.path.join(os.path.dirname(__file__), '..'))
import corpus
n=int(sys.argv[1]) if len(sys.argv)>1 else 8
ratios=[257*k*(256//n) for k in range(n)]
sc=corpus.morph_scene(ratios)
r,stages=corpus.make_product(sc)
r.set_option(2, n)
r.render_batch(stages)
for f in range(n):
    out=r.get_image(frame=f, premultiplied=True).data
    ref=corpus.render_oracle(sc, frame=f)
    print(f, (out!=ref).any(axis=2).sum())
print(r.stats())
