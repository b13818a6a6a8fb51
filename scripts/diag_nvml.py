import time, pynvml, torch, subprocess
pynvml.nvmlInit()
h = pynvml.nvmlDeviceGetHandleByIndex(0)
x = torch.randn(1 << 20, device="cuda")
def t(fn, n=50):
    fn(); t0 = time.perf_counter()
    for _ in range(n): fn()
    return (time.perf_counter() - t0) / n * 1e3
print("clock query ms", t(lambda: pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)))
print("reasons query ms", t(lambda: pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(h)))
print("power query ms", t(lambda: pynvml.nvmlDeviceGetPowerUsage(h)))
# effect on launch+sync latency
def ls():
    y = x * 2; torch.cuda.synchronize()
print("launch+sync ms (quiet)", t(ls, 200))
import threading
stop = False
def loop(period):
    while not stop:
        pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM); pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(h); time.sleep(period)
for period in (0.01, 0.05, 0.2):
    stop = False
    th = threading.Thread(target=loop, args=(period,)); th.start()
    print("launch+sync ms with nvml polling every %.0f ms" % (period * 1e3), t(ls, 400))
    stop = True; th.join()
p = subprocess.Popen(["nvidia-smi", "--query-gpu=clocks.sm,clocks_event_reasons.sw_power_cap", "--format=csv,noheader", "-lms", "200"], stdout=subprocess.DEVNULL)
time.sleep(0.5)
print("launch+sync ms with nvidia-smi -lms 200", t(ls, 400))
p.terminate()
