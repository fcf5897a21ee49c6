import sys, torch, numpy as np
sys.path.insert(0, "/root/repo")
import hcorepp_b200 as hc
sys.argv=['x']
import importlib.util
spec = importlib.util.spec_from_file_location("bench", "/root/repo/bench.py"); b = importlib.util.module_from_spec(spec); spec.loader.exec_module(b)
ctx = hc.RunContext(0); T, nb, acc = 8, 1024, 1e-8
krank = b.rank_for_accuracy(nb, acc); dev = ctx.device; dt = torch.float64
sig = torch.from_numpy(b.spectrum(nb)[:krank].copy()).to(dev)
def synth(n, seed):
    g = torch.Generator(device=dev); g.manual_seed(seed)
    qu,_ = torch.linalg.qr(torch.randn(n, nb, krank, generator=g, dtype=dt, device=dev)); qv,_ = torch.linalg.qr(torch.randn(n, nb, krank, generator=g, dtype=dt, device=dev))
    return qu.transpose(1,2).contiguous(), (qv*sig[None,None,:]).contiguous()
A = hc.TileMatrix(T,T,nb,nb,dt,ctx,compressed=True,max_rank=krank,rank_bound=krank); B = hc.TileMatrix(T,T,nb,nb,dt,ctx,compressed=True,max_rank=krank,rank_bound=krank)
A.load_factors(*synth(T*T,1), krank); B.load_factors(*synth(T*T,2), krank)
C = hc.TileMatrix.zeros_compressed(T,T,nb,nb,dt,ctx)
prm = hc.CompressionParameters(acc)
ctx.stats(reset=True)
for k in range(T):
    hc.tile_matrix_multiplication(A,B,C,1.0,1.0,ctx,prm,k_range=(k,k+1))
    print(k, ctx.stats(reset=True), 'rank', float(C.ranks.float().mean()))
