"""Where does the bf16 error of a UNet call come from?  Runs the fp32 oracle on the GPU with selected operand
classes rounded to bf16/fp16 (fp32 accumulate), and reports eps max-abs-rel against the pure fp32 oracle."""
import sys, os, types
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import unet_oracle as uo

torch.backends.cuda.matmul.allow_tf32 = False
torch.backends.cudnn.allow_tf32 = False


from oracle.precision import unet_forward_rounded


def run(cfg, sd, x, t, c, **kw):
    return unet_forward_rounded(sd, cfg, x, t, c, **kw)


def main():
    which = sys.argv[1] if len(sys.argv) > 1 else "full"
    h = 16 if which == "tiny" else (64 if which == "full64" else 32)   # full64: the mobi_nusc_512 latent
    cfg = uo.tiny_unet_config() if which == "tiny" else uo.default_unet_config(image_size=h)
    sd = {k: v.cuda() for k, v in uo.synth_state_dict(uo.state_dict_shapes(cfg), seed=0).items()}
    inp = uo.synth_inputs(1 if which != "tiny" else 2, h, context_dim=cfg["context_dim"], seed=1, device="cuda")
    x = torch.cat([inp["x_T"], inp["inpaint_image"], inp["inpaint_mask"]], 1)
    x, t, c = torch.cat([x, x]), torch.full((2 * x.shape[0],), 981, device="cuda"), torch.cat([inp["uc"], inp["cond"]])
    ref = run(cfg, sd, x, t, c)
    bf, hf = torch.bfloat16, torch.float16
    cases = {"all bf16": dict(lin=bf, conv=bf, qk=bf, pv=bf), "linear bf16": dict(lin=bf), "conv bf16": dict(conv=bf),
             "qk bf16": dict(qk=bf), "pv bf16": dict(pv=bf), "all fp16": dict(lin=hf, conv=hf, qk=hf, pv=hf),
             "bf16 but attention fp16": dict(lin=bf, conv=bf, qk=hf, pv=hf),
             "bf16 but qk fp16": dict(lin=bf, conv=bf, qk=hf, pv=bf)}
    for name, kw in cases.items():
        out = run(cfg, sd, x, t, c, **kw)
        print("%-28s max-abs-rel %.3e  rel-rms %.3e" % (name, ((out - ref).abs().max() / ref.abs().max()).item(),
                                                        ((out - ref).norm() / ref.norm()).item()))


main()
