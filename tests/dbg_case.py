import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import p3p_cases as cases
from oracle import pillars_oracle as po
from pixelspointspolygons_b200 import PointPillarsEncoder, default_cfg
dev = torch.device("cuda:0")
name = sys.argv[1] if len(sys.argv) > 1 else "x224_then_alias"
tiles, kw = cases.edge_cases()[name]
grid = cases.grid_for(kw)
cfg = default_cfg(device="cuda:0", max_num_points_per_voxel=grid.max_num_points, max_num_voxels=grid.max_voxels, p3p_drop_overflow=grid.drop_overflow)
enc = PointPillarsEncoder(cfg, voxel_encoder={"in_channels": 3, "feat_channels": [64, 384]}, scatter={"in_channels": 384, "output_shape": [28, 28]}).to(dev).eval()
enc.load_state_dict(po.synth_weights(3)[0])
x = torch.nested.nested_tensor([torch.from_numpy(np.ascontiguousarray(t)) for t in tiles], layout=torch.jagged).to(dev)
for prec in ("fp32", "tf32"):
    enc.precision = prec
    for it in range(3):
        out = torch.full((len(tiles), 384, 28, 28), float("nan"), device=dev)
        enc.encode_into(x, out, 0, c_total=384, c_offset=0)
        torch.cuda.synchronize()
        o = out.cpu()
        bad = ~(o == 0).all(1)
        print(prec, it, "nonzero cells:", bad.nonzero().tolist()[:10], "nan:", int(torch.isnan(o).sum()), "absmax", float(o.nan_to_num().abs().max()))
