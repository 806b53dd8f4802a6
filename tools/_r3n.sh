OUT=gpurun_out/r3n; mkdir -p $OUT
timeout 1200 python -m pytest tests/test_gpu_baseline_sizes.py -q --timeout 600 -k "fp32_tensor_core or c5_train" > $OUT/pytest_sel.log 2>&1; echo "pytest rc=$?"; grep -E "^(FAILED|ERROR)|passed|failed|^E  " $OUT/pytest_sel.log | cut -c1-200 | head -40
python - <<'PY'
import sys, torch
sys.path.insert(0,'hybrid-grid-for-hexagonal-and-rectangular-image-processing_b200'); sys.path.insert(0,'.')
from HyGrid import HexFrames as hf
from oracle import hexframes_oracle as HO
torch.manual_seed(0)
for (N,Ci,Co,H,W) in ((8,32,64,64,63),(8,64,128,32,31),(8,32,64,64,64)):
    m=hf.HexConv2d(Ci,Co,0,2,padding=1,bias=False).cuda()
    x=torch.randn(N,Ci,H,W); gy=torch.randn(N,Co,H,W)
    xr=x.clone().requires_grad_(); wr=m.kernel.detach().cpu().requires_grad_()
    ref=HO.hexconv2d(xr,wr,None,0,2,1,1); (ref*gy).sum().backward()
    for mode in ("x3","direct","bf16"):
        m.kernel.grad=None
        xg=x.cuda().requires_grad_()
        hf.set_fp32_tensor_cores(mode=="x3")
        if mode=="bf16":
            with torch.autocast("cuda",dtype=torch.bfloat16): y=m(xg)
        else: y=m(xg)
        (y*gy.cuda()).sum().backward()
        r=lambda a,b: float((a-b).abs().max()/b.abs().max())
        print((N,Ci,Co,H,W),mode,"y",r(y.detach().cpu(),ref.detach()),"dx",r(xg.grad.cpu(),xr.grad),"dw",r(m.kernel.grad.cpu(),wr.grad))
    hf.set_fp32_tensor_cores(True)
PY
