"""PCIe probe + issue-order variants of the end-to-end fragment step (host buffers in, host buffers out).

    python tools/e2e_variants.py
Prints H2D / D2H / simultaneous bandwidth from pinned memory, then ms per e2e step for several issue orders of the
same work (3 levels x [H2D inputs, back_project fwd, D2H volume+count, H2D grad_out, backward, D2H grad_feats]).
"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from deep3dmap_b200 import back_project

dev = torch.device("cuda:0")
t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)


def probe():
    n = 64 << 20
    h1 = torch.empty(n, dtype=torch.uint8).pin_memory(); h2 = torch.empty(n, dtype=torch.uint8).pin_memory()
    d1 = torch.empty(n, dtype=torch.uint8, device=dev); d2 = torch.empty(n, dtype=torch.uint8, device=dev)
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()

    def run(f, reps=10):
        f(); torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(reps):
            f()
        torch.cuda.synchronize()
        return (time.perf_counter() - t0) / reps

    def up():
        with torch.cuda.stream(s1): d1.copy_(h1, non_blocking=True)

    def down():
        with torch.cuda.stream(s2): h2.copy_(d2, non_blocking=True)

    def both():
        up(); down()

    a, b, c = run(up), run(down), run(both)
    print("pcie: H2D %.1f GB/s  D2H %.1f GB/s  both at once %.1f + %.1f GB/s" % (n / a / 1e9, n / b / 1e9, n / c / 1e9, n / c / 1e9))


def cnt_fn(inp):
    return back_project(t(inp["coords"]), t(inp["origin"]), inp["voxel_size"], t(inp["feats"]), t(inp["KRcam"]))[1].cpu().numpy()


def main():
    probe()
    levels = bench.build_fragment_levels(cnt_fn)
    pin = []
    for inp in levels:
        hp = {k: torch.from_numpy(np.ascontiguousarray(inp[k])).pin_memory() for k in ("coords", "origin", "feats", "KRcam", "grad_out")}
        V, B, C, H, W = inp["feats"].shape
        N = inp["coords"].shape[0]
        hp["o_vol"] = torch.empty((N, C + 1), dtype=torch.float32).pin_memory()
        hp["o_cnt"] = torch.empty((N,), dtype=torch.float32).pin_memory()
        hp["o_grad"] = torch.empty((V, B, C, H, W), dtype=torch.float32).pin_memory()
        hp["vs"] = inp["voxel_size"]
        pin.append(hp)
    streams = [torch.cuda.Stream(device=dev) for _ in pin]
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def level_whole(hp):
        c = hp["coords"].to(dev, non_blocking=True); o = hp["origin"].to(dev, non_blocking=True)
        f = hp["feats"].to(dev, non_blocking=True).requires_grad_(True); k = hp["KRcam"].to(dev, non_blocking=True)
        g = hp["grad_out"].to(dev, non_blocking=True)
        vol, cnt = back_project(c, o, hp["vs"], f, k)
        hp["o_vol"].copy_(vol.detach(), non_blocking=True); hp["o_cnt"].copy_(cnt, non_blocking=True)
        vol.backward(g)
        hp["o_grad"].copy_(f.grad, non_blocking=True)

    def v_current():
        main_s = torch.cuda.current_stream()
        for hp, st in reversed(list(zip(pin, streams))):
            st.wait_stream(main_s)
            with torch.cuda.stream(st):
                level_whole(hp)
        for st in streams:
            main_s.wait_stream(st)
        main_s.synchronize()

    def v_small_first():
        main_s = torch.cuda.current_stream()
        for hp, st in zip(pin, streams):
            st.wait_stream(main_s)
            with torch.cuda.stream(st):
                level_whole(hp)
        for st in streams:
            main_s.wait_stream(st)
        main_s.synchronize()

    def phased(order_fwd, order_bwd):
        def f():
            main_s = torch.cuda.current_stream()
            keep = {}
            for i in order_fwd:
                hp, st = pin[i], streams[i]
                st.wait_stream(main_s)
                with torch.cuda.stream(st):
                    c = hp["coords"].to(dev, non_blocking=True); o = hp["origin"].to(dev, non_blocking=True)
                    ft = hp["feats"].to(dev, non_blocking=True).requires_grad_(True); k = hp["KRcam"].to(dev, non_blocking=True)
                    vol, cnt = back_project(c, o, hp["vs"], ft, k)
                    hp["o_vol"].copy_(vol.detach(), non_blocking=True); hp["o_cnt"].copy_(cnt, non_blocking=True)
                    keep[i] = (vol, ft)
            for i in order_bwd:
                hp, st = pin[i], streams[i]
                with torch.cuda.stream(st):
                    g = hp["grad_out"].to(dev, non_blocking=True)
                    vol, ft = keep[i]
                    vol.backward(g)
                    hp["o_grad"].copy_(ft.grad, non_blocking=True)
            for st in streams:
                main_s.wait_stream(st)
            main_s.synchronize()
        return f

    out_streams = [torch.cuda.Stream(device=dev) for _ in pin]

    def phased_split(order_fwd, order_bwd):
        """like `phased`, but every level reads its results back on a second stream, so that the level's next upload
        (grad_out) is not queued behind its own volume read-back"""
        def f():
            main_s = torch.cuda.current_stream()
            keep = {}
            for i in order_fwd:
                hp, st, so = pin[i], streams[i], out_streams[i]
                st.wait_stream(main_s); so.wait_stream(main_s)
                with torch.cuda.stream(st):
                    c = hp["coords"].to(dev, non_blocking=True); o = hp["origin"].to(dev, non_blocking=True)
                    ft = hp["feats"].to(dev, non_blocking=True).requires_grad_(True); k = hp["KRcam"].to(dev, non_blocking=True)
                    vol, cnt = back_project(c, o, hp["vs"], ft, k)
                so.wait_stream(st)
                with torch.cuda.stream(so):
                    hp["o_vol"].copy_(vol.detach(), non_blocking=True); hp["o_cnt"].copy_(cnt, non_blocking=True)
                keep[i] = (vol, ft, cnt)
            for i in order_bwd:
                hp, st, so = pin[i], streams[i], out_streams[i]
                with torch.cuda.stream(st):
                    g = hp["grad_out"].to(dev, non_blocking=True)
                    vol, ft, _ = keep[i]
                    vol.backward(g)
                so.wait_stream(st)
                with torch.cuda.stream(so):
                    hp["o_grad"].copy_(ft.grad, non_blocking=True)
            for st in streams + out_streams:
                main_s.wait_stream(st)
            main_s.synchronize()
        return f

    def single_stream():
        for hp in pin:
            level_whole(hp)
        torch.cuda.current_stream().synchronize()

    variants = [("current (large first, whole level per stream)", v_current),
                ("small first, whole level per stream", v_small_first),
                ("phased fwd 0,1,2 / bwd 2,1,0", phased([0, 1, 2], [2, 1, 0])),
                ("phased fwd 0,1,2 / bwd 0,1,2", phased([0, 1, 2], [0, 1, 2])),
                ("phased fwd 2,1,0 / bwd 2,1,0", phased([2, 1, 0], [2, 1, 0])),
                ("phased 2,1,0 / 2,1,0 + read-back streams", phased_split([2, 1, 0], [2, 1, 0])),
                ("phased 2,1,0 / 0,1,2 + read-back streams", phased_split([2, 1, 0], [0, 1, 2])),
                ("phased fwd 2,1,0 / bwd 2,1,0 (again)", phased([2, 1, 0], [2, 1, 0])),
                ("single stream", single_stream)]
    for name, f in variants:
        for _ in range(3):
            f()
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(20)]
        torch.cuda.synchronize()
        host = 0.0
        for a, b in ev:
            flush.fill_(1)
            a.record()
            t0 = time.perf_counter()
            f()
            host += time.perf_counter() - t0
            b.record()
        torch.cuda.synchronize()
        ms = [a.elapsed_time(b) for a, b in ev]
        print("%-50s  mean %.3f ms  min %.3f ms  (host wall %.3f ms)" % (name, np.mean(ms), np.min(ms), host / 20 * 1e3))
    # host issue cost of the resident eager step
    dl = [dict(coords=t(i["coords"]), origin=t(i["origin"]), vs=i["voxel_size"], feats=t(i["feats"]).requires_grad_(True),
               KR=t(i["KRcam"]), go=t(i["grad_out"])) for i in levels]

    def step():
        for d in dl:
            d["feats"].grad = None
            vol, cnt = back_project(d["coords"], d["origin"], d["vs"], d["feats"], d["KR"])
            vol.backward(d["go"])
    for _ in range(20): step()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(200): step()
    t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
    print("resident eager step: host issue %.1f us/step, with drain %.1f us/step" % ((t1 - t0) / 200 * 1e6, (t2 - t0) / 200 * 1e6))


if __name__ == "__main__":
    main()
