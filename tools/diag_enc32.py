"""margin of the 32-channel encoder on the padded tensor-core path vs the FFMA path (blockpush / buttonpress goldens)"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
from test_gpu_parity import load_case, make_savi, FixedRandn, rel_l2, max_rel, DEV
from playslot_b200 import ops

for case in ["blockpush_sharp", "blockpush", "buttonpress", "sketchy"]:
    try:
        fx, ssd, psd, video, noise, init_noise = load_case(case)
    except Exception as e:
        print(case, "skip", e); continue
    for chans, lin, sa in [((32, 64), "fp16x3", "tc"), ((64,), "fp16x3", "tc"), ((32, 64), "fp32", "fp32"), ((64,), "fp32", "fp32")]:
        savi = make_savi(fx["savi_cfg"], ssd, DEV)
        savi.conv_mode = "fp16x3"
        savi.enc_tc_channels = chans
        ops.set_linear_mode(lin)
        savi.slot_attention.mode = sa
        with torch.no_grad():
            if init_noise is not None:
                with FixedRandn(init_noise):
                    out = savi(video.to(DEV), num_imgs=fx["num_seed"] + fx["num_preds"], decode=False)
            else:
                out = savi(video.to(DEV), num_imgs=fx["num_seed"] + fx["num_preds"], decode=False)
        m = savi.slot_attention.get_attention_masks().cpu(); out = {k: (v.cpu() if torch.is_tensor(v) else v) for k, v in out.items()}
        print(case, chans, lin, sa, "slots max_rel %.3e" % max_rel(out["slot_history"], fx["slot_history"]),
              "masks max_rel %.3e rel_l2 %.3e" % (max_rel(m, fx["sa_masks_last"]), rel_l2(m, fx["sa_masks_last"])))
