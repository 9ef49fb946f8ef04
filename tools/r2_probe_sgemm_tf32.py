"""f32 C -= A B: tcgen05 3xTF32 path (gemm_tf32.cu) against the FP32 FMA kernel and an f64 product."""
import json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lair_b200 import _ffi
L = _ffi.lib(); _ffi.check(L.lair_b200_init(0))
st = torch.cuda.current_stream().cuda_stream
def run(m, n, k, mode, reps=5, pad=0):
    _ffi.set_option("sgemm_tf32", mode)
    g = torch.Generator(device="cuda"); g.manual_seed(m * 7 + n * 3 + k)
    a = torch.rand(m, k + pad, dtype=torch.float32, device="cuda", generator=g) * 2 - 1
    b = torch.rand(k, n + pad, dtype=torch.float32, device="cuda", generator=g) * 2 - 1
    c0 = torch.rand(m, n + pad, dtype=torch.float32, device="cuda", generator=g) * 10
    c = c0.clone()
    _ffi.check(L.lair_b200_sgemm_minus_dev(m, n, k, a.data_ptr(), k + pad, b.data_ptr(), n + pad, c.data_ptr(), n + pad, st))
    torch.cuda.synchronize()
    ref = c0[:, :n].double() - a[:, :k].double() @ b[:, :n].double()
    err = float((c[:, :n].double() - ref).abs().max())
    scale = float(k ** 0.5 * 2.0 ** -24 * 10)
    untouched = bool(torch.equal(c[:, n:], c0[:, n:]))
    ts = []
    for i in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); _ffi.check(L.lair_b200_sgemm_minus_dev(m, n, k, a.data_ptr(), k + pad, b.data_ptr(), n + pad, c.data_ptr(), n + pad, st)); e1.record()
        torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    ms = min(ts)
    print(json.dumps({"bench": "sgemm_minus", "m": m, "n": n, "k": k, "pad": pad, "tf32x3": mode, "max_abs_err": err, "err_over_sqrtk_eps_10": err / scale,
                      "padding_untouched": untouched, "ms": ms, "tflops": 2.0 * m * n * k / ms * 1e-9, "c_gbs": 2.0 * m * n * 4 / ms * 1e-6}), flush=True)
for shape in ((256, 256, 32), (384, 512, 64), (1000, 900, 96), (4096, 4096, 128), (8192, 8192, 128), (16256, 16256, 128), (16128, 16128, 256), (8000, 3000, 64)):
    for mode in (1, 0):
        run(*shape, mode)
run(2048, 2044, 128, 1, pad=4)
run(2048, 2044, 128, 0, pad=4)
