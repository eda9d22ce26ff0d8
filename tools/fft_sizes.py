import torch, time
def t(shape, reps=5):
    x = torch.randn(shape, device="cuda")
    y = torch.fft.rfftn(x); z = torch.fft.irfftn(y, s=shape); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        y = torch.fft.rfftn(x); z = torch.fft.irfftn(y, s=shape)
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps
for shape in [(540,1050,1050),(540,1080,1080),(540,1050,1080),(540,1080,1050),(576,1152,1152),(540,1056,1056),(560,1050,1050),(540,1120,1120),(544,1088,1088),(540,1050,1056),(540,1050,1152)]:
    print(shape, "%.2f ms (R2C+C2R)" % t(shape), "%.2f GB" % (shape[0]*shape[1]*shape[2]*4/1e9))
