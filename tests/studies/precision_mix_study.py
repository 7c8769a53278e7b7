"""CPU study: pose error of selective two-MMA tensor-core operand schemes vs the three-MMA bf16 hi/lo split.
Uses the test oracle with emulated operand roundings; results in profiles/r01e_precision_mix_cpu.json."""
import contextlib, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [ROOT, os.path.join(ROOT, 'oracle')]
import torch, torch.nn.functional as F
import egotap_oracle as orc, weights
from egotap_b200.synthetic import synthetic_heatmaps
def _hi(t): return t.to(torch.bfloat16).float()
def _split(t):
    h=_hi(t); return h,_hi(t-h)
def f16(t): return t.to(torch.float16).float()
def f16split(t):
    h=f16(t); return h,f16(t-h)
@contextlib.contextmanager
def emulate(rule):
    rl, rm, rc = F.linear, torch.matmul, F.conv2d
    def contract(op,a,b,tag):
        mode=rule(tag,a,b)
        if mode=="x3":
            ah,al=_split(a); bh,bl=_split(b); return op(ah,bh)+op(ah,bl)+op(al,bh)
        if mode=="a22_b11":      # A fp16 hi/lo, B fp16: 2 MMAs
            ah,al=f16split(a); b16=f16(b); return op(ah,b16)+op(al,b16)
        if mode=="a11_b22":
            bh,bl=f16split(b); a16=f16(a); return op(a16,bh)+op(a16,bl)
        if mode=="bf16": return op(_hi(a),_hi(b))
        raise ValueError(mode)
    def linear(x,w,b=None):
        y=contract(lambda p,q: rl(p,q),x,w,("linear",tuple(w.shape)))
        return y if b is None else y+b
    def matmul(a,b): return contract(rm,a,b,("matmul",tuple(b.shape[-2:])))
    def conv2d(x,w,b=None,**kw):
        y=contract(lambda p,q: rc(p,q,None,**kw),x,w,("conv",tuple(w.shape)))
        return y if b is None else y+b.view(1,-1,1,1)
    F.linear, torch.matmul, F.conv2d = linear, matmul, conv2d
    try: yield
    finally: F.linear, torch.matmul, F.conv2d = rl, rm, rc
res={}
for preset in ("UnrealEgo","EgoCap"):
    sd=weights.make_state_dict(preset,seed=5)
    x=synthetic_heatmaps(preset,4,seed=1234,kind="gauss")
    with torch.no_grad():
        truth=orc.forward({k:v.double() for k,v in sd.items()},x.double(),preset)
        def run(rule):
            with emulate(rule): return orc.parity_report(orc.forward(sd,x,preset),truth)["rel"]
        mlp=lambda tag: tag[0]=="linear" and tag[1] in ((4096,1024),(1024,4096))
        qkvo=lambda tag: tag[0]=="linear" and tag[1]==(1024,1024)
        cases={
          "all x3": lambda t,a,b:"x3",
          "mlp a22_b11": lambda t,a,b:"a22_b11" if mlp(t) else "x3",
          "mlp a11_b22": lambda t,a,b:"a11_b22" if mlp(t) else "x3",
          "mlp-down only a22_b11": lambda t,a,b:"a22_b11" if (t[0]=="linear" and t[1]==(1024,4096)) else "x3",
          "mlp-up only a22_b11": lambda t,a,b:"a22_b11" if (t[0]=="linear" and t[1]==(4096,1024)) else "x3",
          "qkvo a22_b11": lambda t,a,b:"a22_b11" if qkvo(t) else "x3",
          "all vit linear a22_b11": lambda t,a,b:"a22_b11" if (mlp(t) or qkvo(t)) else "x3",
          "all a22_b11": lambda t,a,b:"a22_b11",
        }
        for k,r in cases.items():
            print(preset,k,"%.2e"%run(r),flush=True)
