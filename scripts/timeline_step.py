"""Kernel timeline of ONE CUDA-graph replay of the bench step (torch.profiler / CUPTI): per-stream busy time, idle gaps of the
device as a whole, and which kernels the gaps follow.  Writes a text summary to stdout."""
import os, sys, collections, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from stylegan_v_b200.synthesis import SynthesisNetwork
from torch.profiler import profile, ProfilerActivity
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
dev = torch.device('cuda')
N = 32
net = SynthesisNetwork(img_resolution=256).to(dev).train()
ws = torch.randn(N, net.num_ws, net.w_dim, device=dev)
t = torch.zeros(N, 1, device=dev)
mz = torch.randn(N, net.motion_encoder.traj_len(), 512, device=dev)
dimg = torch.randn(N, 3, 256, 256, device=dev)


def step():
    for p in net.parameters():
        p.grad = None
    w = ws.clone().requires_grad_(True)
    img = net(w, t, motion_z=mz)
    (img * dimg).sum().backward()


side = torch.cuda.Stream()
side.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(side):
    for _ in range(2):
        step()
torch.cuda.current_stream().wait_stream(side)
torch.cuda.synchronize()
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g):
    step()
for _ in range(3):
    g.replay()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    g.replay()
    torch.cuda.synchronize()
evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA and e.time_range is not None]
ks = []
for e in evs:
    name = e.name
    if name.startswith('Memcpy') or name.startswith('Memset') or 'graph' in name.lower():
        kind = 'mem'
    else:
        kind = 'k'
    ks.append((e.time_range.start, e.time_range.end, name, getattr(e, 'stream', None) if hasattr(e, 'stream') else None, kind))
ks.sort()
if not ks:
    print('no CUDA events captured'); sys.exit(0)
t0, t1 = ks[0][0], max(k[1] for k in ks)
print(f'{len(ks)} device activities, span {(t1 - t0) / 1e3:.3f} ms')
# union busy / gaps
gaps = []
cur_end = ks[0][1]; last_name = ks[0][2]
busy = ks[0][1] - ks[0][0]
for s, e, name, stream, kind in ks[1:]:
    if s > cur_end:
        gaps.append((s - cur_end, last_name, name))
        busy += e - s
        cur_end = e; last_name = name
    else:
        if e > cur_end:
            busy += e - cur_end
            cur_end = e; last_name = name
print(f'device busy (union of all streams) {busy / 1e3:.3f} ms, idle {(t1 - t0 - busy) / 1e3:.3f} ms in {len(gaps)} gaps')
# overlap: sum of durations vs union
tot = sum(e - s for s, e, *_ in ks)
print(f'sum of activity durations {tot / 1e3:.3f} ms  (overlap = {(tot - busy) / 1e3:.3f} ms)')
by = collections.defaultdict(float); cnt = collections.Counter()
for s, e, name, stream, kind in ks:
    short = name.split('(')[0][:70]
    by[short] += e - s; cnt[short] += 1
print('top kernels by total time:')
for k, v in sorted(by.items(), key=lambda kv: -kv[1])[:22]:
    print(f'  {v / 1e3:8.3f} ms x{cnt[k]:<4d} {k}')
gb = collections.defaultdict(float); gc = collections.Counter()
for d, a, b in gaps:
    gb[a.split('(')[0][:60]] += d; gc[a.split('(')[0][:60]] += 1
print('idle time grouped by the kernel that PRECEDES the gap:')
for k, v in sorted(gb.items(), key=lambda kv: -kv[1])[:15]:
    print(f'  {v / 1e3:8.3f} ms x{gc[k]:<4d} after {k}')
hist = collections.Counter()
for d, a, b in gaps:
    hist[min(int(d // 1000), 20)] += 1      # microsecond buckets
print('gap length histogram (us: count):', dict(sorted(hist.items())))
# ---- where the step is SERIAL: time during which exactly one activity is in flight, grouped by that activity; and the tail of the step ----
edges = []
for i, (s, e, name, stream, kind) in enumerate(ks):
    edges.append((s, 1, i)); edges.append((e, -1, i))
edges.sort()
live = set(); alone = collections.defaultdict(float); multi = 0.0; prev = edges[0][0]
for tt, d, i in edges:
    if tt > prev and live:
        if len(live) == 1:
            alone[ks[next(iter(live))][2].split('(')[0][:70]] += tt - prev
        else:
            multi += tt - prev
    prev = tt
    (live.add if d > 0 else live.discard)(i)
print(f'time with >= 2 activities in flight {multi / 1e3:.3f} ms; time with exactly one, by kernel:')
for k, v in sorted(alone.items(), key=lambda kv: -kv[1])[:25]:
    print(f'  {v / 1e3:8.3f} ms  {k}')
streams = collections.defaultdict(float)
for s, e, name, stream, kind in ks:
    streams[stream] += e - s
print('busy time per stream:', {str(k): round(v / 1e3, 3) for k, v in streams.items()})
print('last 60 activities (start offset from the end of the step in ms, duration ms, stream, name):')
for s, e, name, stream, kind in sorted(ks, key=lambda k: k[1])[-60:]:
    print(f'  {(s - t1) / 1e3:9.3f} {(e - s) / 1e3:8.3f}  {stream}  {name.split("(")[0][:80]}')
