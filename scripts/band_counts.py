"""Development aid: render row bands of a 1920x1080 demo frame separately (for ncu per-launch instruction counts)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from terminalraytracer_b200 import renderer as R, scene as S
w, h = 1920, 1080
sky = S.get_skybox("milky_way")
rd = R.Renderer(0, sky)
sc = S.SceneData(w, h, sky).set_time(3.7)
rd.set_scene(sc)
q = torch.empty(w * h * 4, dtype=torch.uint8, device="cuda")
for (r0, r1) in [(0, 1080), (0, 1080), (960, 1080), (0, 120), (480, 600)]:
    rd.render_rows_quant(w, h, r0, r1, q.data_ptr())
    rd.synchronize()
    print("band", r0, r1)
rd.close()
