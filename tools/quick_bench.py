import time, torch, sys
sys.path.insert(0, "/root/repo")
import audiodeepfake_detection_b200 as afd
from audiodeepfake_detection_b200.wavelets import Wavelet
def t(fn, n=20, w=5):
    for _ in range(w): fn()
    torch.cuda.synchronize(); e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize(); return e0.elapsed_time(e1)/n
B=4096
x=torch.randn(B,22050,device="cuda")*0.1
for name in ["coif4","sym5","db2","haar"]:
    w=Wavelet(name)
    ms=t(lambda: afd.wavelet_packet_features(x,w,8,log_scale=True))
    print(name, "B",B, "ms",ms, "frames/s", B/ms*1e3)
xs=x[:128].contiguous()
for name in ["coif4","sym5"]:
    w=Wavelet(name); ms=t(lambda: afd.wavelet_packet_features(xs,w,8,log_scale=True)); print(name,"B128 ms",ms,"frames/s",128/ms*1e3)
