import torch


def report(name, got, ref, rtol, atol):
    got = got.float()
    ref = ref.float()
    err = (got - ref).abs()
    tol = atol + rtol * ref.abs()
    bad = err > tol
    nbad = int(bad.sum())
    msg = (f"{name}: max_abs_err={err.max().item():.3e} mean_abs_err={err.mean().item():.3e} "
           f"ref_abs_mean={ref.abs().mean().item():.3e} bad={nbad}/{err.numel()} (rtol={rtol}, atol={atol})")
    print(msg)
    if nbad:
        idx = bad.nonzero()[:8].tolist()
        print("  first bad idx:", idx)
        for i in idx[:4]:
            print("   ", i, float(got[tuple(i)]), float(ref[tuple(i)]))
        if got.dim() == 2:
            r, c = got.shape
            rb = bad.any(dim=1).nonzero().flatten()
            cb = bad.any(dim=0).nonzero().flatten()
            print(f"  bad rows: {rb.numel()}/{r} first {rb[:16].tolist()}; bad cols: {cb.numel()}/{c} first {cb[:16].tolist()}")
    return nbad == 0, msg


def check(name, got, ref, rtol, atol):
    ok, msg = report(name, got, ref, rtol, atol)
    assert ok, msg
