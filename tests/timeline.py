"""DUALAR_TIMELINE=1 python tests/timeline.py : in-kernel %globaltimer stamps of one decode step (CTA 0).

Persistent kernel (default): slot 1+ph holds {phase start, input staged, compute done} of phase ph; slot 0 = kernel
entry / exit.  DUALAR_MEGA=0: one slot per kernel of the step graph (entry, dependency wait returned, prologue done, exit)."""
import collections
import os
import sys
from pathlib import Path

import torch

os.environ["DUALAR_TIMELINE"] = "1"
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from fish_tts_b200.config import s1_mini_config  # noqa: E402
from fish_tts_b200.engine import DualAREngine  # noqa: E402
from fish_tts_b200.synthetic import make_state_dict, synthetic_prompt  # noqa: E402

cfg = s1_mini_config()
eng = DualAREngine(cfg, make_state_dict(cfg, seed=0), device=0, seed=1234)
n_step, n_pre = eng.launches_per_step()
prompt = synthetic_prompt(cfg, 3, 215, 5, seed=1)
eng.prefill(prompt, 600, temperature=0.7, top_p=0.8, repetition_penalty=1.1)
eng.decode(300)
torch.cuda.synchronize()
L = cfg.n_layer

if n_step == 1:
    names = []
    for i in range(L):
        names += [f"S.qkv", "S.attn", "S.merge", "S.wo", "S.w13", "S.w2"]
    for l in range(cfg.n_fast_layer):      # fast pass 0, one slice of the LM head in front of every phase
        names += ["H.head", "F.qkv", "H.head", "F.wo", "H.head", "F.w13", "H.head", "F.w2"]
    names += ["H.stat", "H.cand"]
    table = os.environ.get("DUALAR_T0", "1") != "0"      # first-layer wqkv of passes >= 1 comes from the code table: no phase
    for p in range(1, cfg.num_codebooks):
        for l in range(cfg.n_fast_layer):
            names += (["F.wo", "F.w13", "F.w2"] if (table and l == 0) else ["F.qkv", "F.wo", "F.w13", "F.w2"])
        names.append("F.head")
    nph = len(names)
    full = eng.read("timeline").numpy().astype("int64")
    dbg = full[512:512 + nph, :4]
    tl = full[: nph + 1]
    t0 = tl[0, 0]
    start = (tl[1:, 0] - t0) / 1e3
    staged = (tl[1:, 1] - t0) / 1e3
    done = (tl[1:, 2] - t0) / 1e3
    end = (tl[0, 3] - t0) / 1e3
    print(f"step total (kernel entry -> exit, CTA 0): {end:.1f} us, {nph} phases")
    agg = collections.defaultdict(lambda: [0, 0.0, 0.0, 0.0])
    pfirst = (tl[1:, 4] - t0) / 1e3
    plast = (tl[1:, 5] - t0) / 1e3
    waits = tl[1:, 6] / 1e3
    print("  ph name      start   +staged  +compute  period | ring wait (in compute) | producer: began / finished issuing this phase, relative to the phase start")
    for i, nm in enumerate(names):
        nxt = start[i + 1] if i + 1 < nph else end
        period = nxt - start[i]
        st = staged[i] - start[i] if tl[1 + i, 1] else float("nan")
        cp = done[i] - (staged[i] if tl[1 + i, 1] else start[i])
        a = agg[nm]; a[0] += 1; a[1] += 0.0 if st != st else st; a[2] += cp; a[3] += period
        if i < 14 or 6 * L - 7 <= i < 6 * L + 40 or i >= nph - 12:
            print(f"{i:4d} {nm:8s} {start[i]:8.2f} {st:8.2f} {cp:8.2f} {period:8.2f} | {waits[i]:6.2f} | {pfirst[i] - start[i]:8.2f} {plast[i] - start[i]:8.2f} | cyc take {dbg[i,0]:6d} mma {dbg[i,1]:6d} part {dbg[i,2]:6d} fold {dbg[i,3]:6d}")
    print("kind       n   staged  compute   period      sum")
    for k, a in agg.items():
        n = a[0]
        print(f"{k:8s} {n:4d} {a[1] / n:8.2f} {a[2] / n:8.2f} {a[3] / n:8.2f} {a[3]:8.1f}")
    # warp 0's first unit of every GEMV phase: staged -> tile landed -> MMA chunk done -> partials handed in -> own loop done
    un = full[512:512 + nph, 4:7]
    uacc = collections.defaultdict(lambda: [0, 0.0, 0.0, 0.0, 0.0])
    for i, nm in enumerate(names):
        if nm.split(".")[1] in ("qkv", "wo", "w13", "w2", "head") and nm != "H.head" and un[i, 0] and un[i, 1] and un[i, 2] and tl[1 + i, 1]:
            u = uacc[nm]; u[0] += 1
            u[1] += (un[i, 0] - tl[1 + i, 1]) / 1e3; u[2] += (un[i, 1] - un[i, 0]) / 1e3; u[3] += (un[i, 2] - un[i, 1]) / 1e3; u[4] += (tl[1 + i, 2] - un[i, 2]) / 1e3
    print("  warp 0, first unit (mean us): kind   staged->landed | mma chunk | partials + hand-in | rest of its loop (fold if last, more units)")
    for k, u in uacc.items():
        n = u[0]
        print(f"    {k:8s} {u[1] / n:8.2f} {u[2] / n:8.2f} {u[3] / n:8.2f} {u[4] / n:8.2f}")
    fa = full[512:512 + nph, :6]
    acc = [0.0] * 5; nfa = 0
    for i, nm in enumerate(names):
        if nm == "F.wo" and fa[i, 0]:
            st_ = tl[1 + i, 0]; nfa += 1
            for k_, (x0, x1) in enumerate([(st_, fa[i, 0]), (fa[i, 0], fa[i, 1]), (fa[i, 1], fa[i, 2]), (fa[i, 2], fa[i, 3]), (fa[i, 3], tl[1 + i, 1])]):
                acc[k_] += (x1 - x0) / 1e3
    if nfa:
        print("  F.wo staging (mean us): poll %.2f | barrier %.2f | rope %.2f | kv+scores %.2f | softmax+PV+store+barrier %.2f" % tuple(x / nfa for x in acc))
    acc = [0.0] * 4; nsa = 0
    for i, nm in enumerate(names):
        if nm == "S.attn" and fa[i, 0] and i > 6:
            nsa += 1
            for k_, (x0, x1) in enumerate([(tl[1 + i, 0], fa[i, 0]), (fa[i, 0], fa[i, 1]), (fa[i, 1], fa[i, 2]), (fa[i, 2], tl[1 + i, 2])]):
                acc[k_] += (x1 - x0) / 1e3
    if nsa:
        print("  S.attn (CTA 0, mean us): poll q/k/v %.2f | barrier+norm+rope+barrier %.2f | tile walk %.2f | warp merge + publish %.2f" % tuple(x / nsa for x in acc))
        w0 = [ (fa[i,3]-fa[i,1])/1e3 for i,nm in enumerate(names) if nm == 'S.attn' and fa[i,0] and i > 6]
        w1 = [ (fa[i,4]-fa[i,3])/1e3 for i,nm in enumerate(names) if nm == 'S.attn' and fa[i,0] and i > 6]
        w2 = [ (fa[i,2]-fa[i,4])/1e3 for i,nm in enumerate(names) if nm == 'S.attn' and fa[i,0] and i > 6]
        print('    tile walk split: init+wait for tile %.2f | warp 0 compute %.2f | barrier after tile + state %.2f' % (sum(w0)/len(w0), sum(w1)/len(w1), sum(w2)/len(w2)))
    hs = full[1024:1024 + nph, :4]
    i_hc = names.index("H.cand")
    if hs[i_hc, 0]:
        x = [tl[1 + i_hc, 0]] + [hs[i_hc, k] for k in range(4)] + [tl[1 + i_hc, 2]]
        print("  H.cand (CTA 0, us): poll stats + prefix %.2f | write candidates %.2f | fetch candidates %.2f | sampler %.2f | embedding publish %.2f" % tuple((x[k + 1] - x[k]) / 1e3 for k in range(5)))
    for i, nm in enumerate(names):
        if nm == "F.head" and i < nph - 1:
            print(f"  F.head ph {i}: start {start[i]:.2f} staged +{staged[i]-start[i]:.2f} compute +{done[i]-staged[i]:.2f} | polled +{(hs[i,0]-t0)/1e3-done[i]:.2f} | m,S +{(hs[i,1]-hs[i,0])/1e3:.2f} | sampler +{(hs[i,2]-hs[i,1])/1e3:.2f} | publish+bar +{(hs[i,3]-hs[i,2])/1e3:.2f}")
    raw2 = eng.read("timeline2").numpy().astype("int64").reshape(-1)
    t2 = raw2[: nph * 148 * 2].reshape(nph, 148, 2)
    pw = raw2[400 * 160 * 2: 400 * 160 * 2 + nph * 148].reshape(nph, 148) / 1e3      # ring wait of thread 0 of each CTA in each phase, us
    import numpy as np
    print("per-CTA view (GEMV phases): staged = all inputs seen, done = own units stored; us relative to the phase's earliest 'staged'")
    print("  ph name     first_staged last_staged | first_done last_done (cta) | CTA0 done | next phase first_staged")
    for i, nm in enumerate(names):
        if t2[i, :, 0].min() == 0: continue
        if not (12 <= i < 24 or 6 * L <= i < 6 * L + 12): continue
        s0 = t2[i, :, 0].min()
        nxt = [j for j in range(i + 1, nph) if t2[j, :, 0].min() > 0]
        nx = (t2[nxt[0], :, 0].min() - s0) / 1e3 if nxt else float('nan')
        print(f"{i:4d} {nm:8s} {0.0:8.2f} {(t2[i,:,0].max()-s0)/1e3:8.2f} (cta {int(t2[i,:,0].argmax()):3d}) | {(t2[i,:,1].min()-s0)/1e3:8.2f} {(t2[i,:,1].max()-s0)/1e3:8.2f} (cta {int(t2[i,:,1].argmax()):3d}) | {(t2[i,0,1]-s0)/1e3:8.2f} | {nx:8.2f} | ring wait mean {pw[i].mean():.2f} max {pw[i].max():.2f} (cta {int(pw[i].argmax())})")
    gem = [i for i in range(nph) if t2[i, :, 0].min() > 0]
    print(f"sum over GEMV phases of (max over CTAs of thread-0 ring wait): {sum(pw[i].max() for i in gem):.1f} us; of the mean: {sum(pw[i].mean() for i in gem):.1f} us; of (last done - first staged): {sum((t2[i,:,1].max() - t2[i,:,0].min()) / 1e3 for i in gem):.1f} us; of (last staged - first staged): {sum((t2[i,:,0].max() - t2[i,:,0].min()) / 1e3 for i in gem):.1f} us")
    wc = raw2[400 * 148 * 2: 400 * 148 * 2 + 148 * 16].reshape(148, 16) / 1965.0
    print(f"ring waits per warp over the whole step (us): mean {wc.mean():.1f}  min {wc.min():.1f}  max {wc.max():.1f}; per-CTA mean: min {wc.mean(1).min():.1f} (cta {int(wc.mean(1).argmin())})  max {wc.mean(1).max():.1f} (cta {int(wc.mean(1).argmax())}); CTA 0: {wc[0].mean():.1f}")
    print('slow-head candidates of the last step:', int(eng.read('n_cand')[0]), ' nucleus sizes (slow, fast heads):', eng.read('nucleus').tolist())
    slow_end = start[6 * L]
    i_hs = names.index("H.stat")
    print(f"slow stack {slow_end:.1f} us | LM head + fast pass 0 {start[i_hs] - slow_end:.1f} us | slow sampler {start[i_hs + 2] - start[i_hs]:.1f} us | fast passes 1.. {end - start[i_hs + 2]:.1f} us")
    sys.exit(0)

tl = eng.read("timeline")[:n_step].numpy()
g = tl[:, :4].astype("int64")
t0 = g[0, 0]
names = ["embed"]
for i in range(L):
    names += [f"L{i}.qkv", f"L{i}.attn", f"L{i}.wo", f"L{i}.w13", f"L{i}.w2"]
names += ["head", "select"]
for p in range(cfg.num_codebooks):
    for l in range(cfg.n_fast_layer):
        names += [f"F{p}.{l}.qkv", f"F{p}.{l}.wo", f"F{p}.{l}.w13", f"F{p}.{l}.w2"]
    if p:
        names.append(f"F{p}.head")
assert len(names) == n_step, (len(names), n_step)
print("slot name           entry    wait_ret  pro_done  end(b0)   | gap_from_prev_end  wait-entry  pro-wait  end-pro   (us, relative to step start)")
prev_end = None
agg = collections.defaultdict(lambda: [0, 0.0, 0.0, 0.0, 0.0, 0.0])
for i, nm in enumerate(names):
    e, w, p, x = [(v - t0) / 1e3 for v in g[i]]
    kind = nm.split(".")[-1] if "." in nm else nm
    kind = ("F." if nm.startswith("F") else "S.") + kind
    a = agg[kind]; a[0] += 1; a[1] += w - e; a[2] += p - w; a[3] += x - p; a[4] += (x - e)
    if i + 1 < n_step:
        a[5] += (g[i + 1, 1] - g[i, 1]) / 1e3     # wait-return to next wait-return = the serial period
    if i < 14 or (145 <= i < 170) or i > n_step - 20:
        print(f"{i:4d} {nm:12s} {e:9.2f} {w:9.2f} {p:9.2f} {x:9.2f}   | {'' if prev_end is None else f'{e - prev_end:8.2f}':>8s} {w - e:10.2f} {p - w:9.2f} {x - p:8.2f}")
    prev_end = x
print(f"step total (first entry -> last end): {(g[-1, 3] - t0) / 1e3:.1f} us")
print("kind        n   wait-entry  pro-wait  end-pro  total(b0)  serial period (wait_ret -> next wait_ret)")
for k, a in agg.items():
    n = a[0]
    print(f"{k:10s} {n:3d} {a[1] / n:9.2f} {a[2] / n:9.2f} {a[3] / n:8.2f} {a[4] / n:9.2f} {a[5] / n:9.2f}   sum {a[5]:8.1f}")
