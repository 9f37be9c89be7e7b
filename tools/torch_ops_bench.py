"""Times the torch ops the exchange layer (omega_h_b200/dist.py) leans on, at pass-3 sizes."""
import time, torch
dev = torch.device("cuda")
n = 112_000_000
def t(name, f, reps=5):
    f(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps): r = f()
    torch.cuda.synchronize()
    print("%-40s %8.3f ms" % (name, (time.perf_counter() - t0) * 1e3 / reps), flush=True)
    return r
key = torch.arange(n, device=dev, dtype=torch.int64)
cnt = torch.ones(n, device=dev, dtype=torch.int64)
rk = torch.zeros(n, device=dev, dtype=torch.int32); rk[::1000] = 1
dp = torch.zeros(n, device=dev, dtype=torch.int8)
counted = t("rk == me", lambda: rk == 0)
w = t("where(counted, cnt, 0)", lambda: torch.where(counted, cnt, 0))
incl = t("cumsum int64", lambda: torch.cumsum(w, 0))
pre = t("incl - w", lambda: incl - w)
def mkstart():
    s = counted.clone()
    s[1:] &= ~(counted[:-1] & (key[:-1] + 1 == key[1:]))
    return s
start = t("start flags", mkstart)
rid = t("cumsum bool", lambda: torch.cumsum(start, 0) - 1)
rf = t("nonzero(start) sparse", lambda: torch.nonzero(start).flatten())
print("runs", rf.numel())
t("nonzero(counted) dense", lambda: torch.nonzero(counted).flatten())
rb = torch.zeros(rf.numel(), device=dev, dtype=torch.int64)
t("where(counted, rb[rid]+pre, pre)", lambda: torch.where(counted, rb[rid.clamp(min=0)] + pre, pre))
t("want nonzero", lambda: torch.nonzero((~counted) & (dp <= 1)).flatten())
t("cat 4", lambda: torch.cat([cnt[:4_000_000], cnt[:30_000_000], cnt[:52_000_000], cnt[:26_000_000]]))
off = torch.arange(52_000_001, device=dev, dtype=torch.int32)
t("off diff to int64 (52M)", lambda: (off[1:] - off[:-1]).to(torch.int64))
t("slice - const (52M)", lambda: pre[:52_000_000] - 5)
t("key slice += (52M)", lambda: key[:52_000_000].add_(1))
t("empty+copy int64 112M", lambda: key.clone())
e = torch.randint(0, 30_000_000, (300_000,), device=dev)
t("gather small", lambda: key[e])
t("searchsorted 300K in 1M", lambda: torch.searchsorted(key[:1_000_000], e))
t("tolist 2", lambda: key[:2].tolist(), reps=20)
t("item", lambda: key[5].item(), reps=20)
