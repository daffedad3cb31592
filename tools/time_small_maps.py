import sys, os, torch
sys.path.insert(0, os.getcwd())
import cl4wsis_b200 as cl4
def t(B,C,H,W,dil,path):
    if path: os.environ["CL4_SWEEP"]=path
    else: os.environ.pop("CL4_SWEEP",None)
    x=torch.rand(B,3,H,W,device="cuda"); m=torch.rand(B,C,H,W,device="cuda").softmax(1)
    mod=cl4.PAMR(10,dil).cuda()
    for _ in range(3): mod(x,m)
    torch.cuda.synchronize()
    e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): mod(x,m)
    e1.record(); torch.cuda.synchronize()
    print(f"B{B} C{C} {H}x{W} D{len(dil)} path={path or 'fused':5s} {e0.elapsed_time(e1)/20*1000:9.1f} us/call")
for cfg in [(16,21,32,32,[1,2,4,8,12]),(24,21,32,32,[1,2,4,8,12]),(16,81,56,56,[1,2,4,8,12]),(16,21,64,64,[1,2,4,8,12,24])]:
    for path in (None,"tma","v1"):
        t(*cfg,path)
