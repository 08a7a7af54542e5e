"""CPU experiment (uses the oracle; not part of the product): how much end-to-end error would cheaper SiLU formulations
add on top of fp16 operand / fp16 activation storage?  Simulates the engine's numerics with torch on the CPU."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch, torch.nn.functional as F
from oracle import yolov5 as oy, letterbox as olb
from megadetector_b200 import checkpoint
from megadetector_b200.synth import synth_image

torch.set_num_threads(32)
weights, _arch_dict = checkpoint.load_model('synthetic://yolov5x6?seed=0')
w = {k: (v if torch.is_tensor(v) else torch.from_numpy(v)).float() for k, v in weights.items()}
img = synth_image(3)
arch = oy.Arch()
lb, _, _ = olb.letterbox(img, 1280, 64, True, True)
x = torch.from_numpy(olb.to_model_input(lb))[None]

def h(t): return t.half().float()
MODE = 'exact'
def silu_var(y):
    if MODE == 'exact':
        return y * torch.sigmoid(y)
    if MODE == 'ex2_f16':   # e = 2^t evaluated in half precision (input and output rounded to fp16), 1/(1+e) in fp32
        t = h(torch.clamp(y * -1.4426950408889634, max=15.0))
        e = h(torch.exp2(t))
        return y / (1.0 + e)
    if MODE == 'tanh11':    # tanh with relative error 2^-11 (modelled as rounding to 11 significant bits)
        hh = 0.5 * y
        t = torch.tanh(hh)
        m, ex = torch.frexp(t)
        t = torch.ldexp(torch.round(m * 2048) / 2048, ex)
        return hh * t + hh
def conv_sim(xx, ww, name, s=1, p=0, act=True, trace=None):
    y = F.conv2d(h(xx), h(ww[name + '.weight']), ww[name + '.bias'], stride=s, padding=p)
    if act:
        y = silu_var(y)
    return h(y)   # activations are stored in fp16 (the residual add of the engine happens before this rounding too)
ref = oy.forward(x, w, arch)[0].numpy()
orig = oy._conv
oy._conv = conv_sim
res = {}
for MODE in ['exact', 'ex2_f16', 'tanh11']:
    out = oy.forward(x, w, arch)[0].numpy()
    d = np.abs(out - ref)
    print('%-8s max score err %.4f  mean score err %.5f  max box err %.2f px  mean box err %.4f' % (MODE, d[:, 4:].max(), d[:, 4:].mean(), d[:, :4].max(), d[:, :4].mean()), flush=True)
