import numpy as np, torch, time, sys
sys.path.insert(0,".")
from image_segmenter_b200.engine import get_engine
e=get_engine(0)
img=np.random.default_rng(0).integers(0,256,(8192,8192,4),dtype=np.uint8)
fn=lambda: e.upload_rgba(img)
fn(); torch.cuda.synchronize()
t=time.perf_counter()
for _ in range(5): d=fn()
torch.cuda.synchronize()
dt=(time.perf_counter()-t)/5
print(round(dt*1e3,2),"ms", round(img.nbytes/dt/1e9,1),"GB/s")
