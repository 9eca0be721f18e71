"""Development aid for ncu: three K1 launches of the bench frame (7680x4320 demo scene) on the library TRT_B200_LIB selects."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from terminalraytracer_b200 import renderer as R, scene as S
w, h = int(os.environ.get("W", 7680)), int(os.environ.get("H", 4320))
sky = S.get_skybox("milky_way")
rd = R.Renderer(0, sky)
rd.use_stream(torch.cuda.current_stream().cuda_stream)
sc = S.SceneData(w, h, sky, kind=os.environ.get("KIND", "demo")).set_time(3.7)
rd.set_scene(sc)
quant = torch.zeros(w * h * 4, dtype=torch.uint8, device="cuda")
for i in range(3):
    rd.render_rows_quant(w, h, 0, h, quant.data_ptr())
torch.cuda.synchronize()
rd.use_stream(None)
rd.close()
