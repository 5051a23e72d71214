import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pyca_b200
from tools.quick_bench import timeit, pd
d3 = pd("params_lenia_abominable_mongolic.npz")
d1 = {k: (v[:, :1, :1].clone() if isinstance(v, torch.Tensor) else v) for k, v in d3.items()}
SIZES = [tuple(int(v) for v in a.split('x')) for a in sys.argv[1:]] or [(1024, 1024), (4096, 3976), (16384, 16384)]
for (H, W) in SIZES:
    a = pyca_b200.Lenia((H, W), num_channels=1, params=d1, state_init=torch.rand(1, 1, H, W), device="cuda", conv_path="fft")
    a.step()
    fs = a._fs
    from pyca_b200 import _lib
    fn = _lib.lib().b200ca_lenia_step_fft
    args = (_lib.ptr(fs.buf), _lib.ptr(fs.other), _lib.ptr(a._kfft), _lib.ptr(a.mu), _lib.ptr(a.sigma), _lib.ptr(a.weights),
            _lib.ptr(a._fft_ws), 1, 1, 1, H, W, 0.1, 0, 1, torch.cuda.current_stream().cuda_stream)
    ms = timeit(lambda: fn(*args), 200)
    print(f"only={os.environ.get('B200CA_FFT_ONLY','0')} {H}x{W}: {ms*1e3:.1f} us per call (bare ctypes call, no python wrapper)")
