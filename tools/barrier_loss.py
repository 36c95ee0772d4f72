"""Where does k_mp_frames<32> wait?  Replays the CTA composition of one aged env-step from the per-env per-frame
busy cycles (mpb_frame_profile) and the schedule that step used: for every barrier unit (a frame, or the
free-running map-scan block 1..8) the CTA pays the MAX over its 32 envs, an env alone would pay its own cycles.
Prints, per phase that ends a unit, the share of the CTA time and the max/mean ratio, and the per-env-frame
cycle quantiles per phase.

  python tools/barrier_loss.py [envs] [age] [--load gpurun_out/frame_prof.npz]
"""
import sys
import numpy as np


def collect(n, age):
    import torch
    sys.path.insert(0, '.')
    from gymcity_b200.micropolis_env import MicropolisEnv
    env = MicropolisEnv(num_envs=n); env.seed(0); env.setMapSize(16, max_step=10**9, empty_start=False); env.reset()
    g = torch.Generator(device='cuda'); g.manual_seed(0)
    for t in range(age):
        env.step(torch.randint(0, 20 * 256, (n,), device='cuda', generator=g))
    eng = env.micro.engine
    eng.frame_profile(True)
    profs, perms = [], []
    for k in range(4):                                   # four consecutive steps: how well does a step predict the next?
        env.step(torch.randint(0, 20 * 256, (n,), device='cuda', generator=g))
        profs.append(eng.frame_profile(True, read=True)[:, :200].astype(np.uint32))
        perm, G = eng.read_schedule()                    # the order this step used (rebuilt at the start of every step)
        perms.append(perm)
    prof, perm = profs[0], perms[0]
    np.savez_compressed('gpurun_out/frame_prof.npz', prof=prof, perm=perm, groups=G, age=age, more=np.stack(profs[1:]),
                        perms=np.stack(perms))
    return prof, perm, G, age


def main():
    args = [a for a in sys.argv[1:] if not a.startswith('--')]
    if '--load' in sys.argv:
        z = np.load(sys.argv[sys.argv.index('--load') + 1])
        prof, perm, G, age = z['prof'], z['perm'], int(z['groups']), int(z['age'])
    else:
        n = int(args[0]) if args else 32768
        age = int(args[1]) if len(args) > 1 else 40
        prof, perm, G, age = collect(n, age)
    n = prof.shape[0]
    c = prof.astype(np.float64)
    ph_of = (np.arange(200) + ((age + 1) * 200) % 16) % 16
    heavy = n // 64 if G >= 3 else 0
    print("envs %d, groups %d, heavy %d; busy cycles per env-step: mean %.3g p50 %.3g p99 %.3g max %.3g"
          % (n, G, heavy, c.sum(1).mean(), np.median(c.sum(1)), np.percentile(c.sum(1), 99), c.sum(1).max()))
    print("per env-frame busy cycles by phase:  mean / p50 / p90 / p99 / p99.9 / max, share of all busy cycles")
    tot = c.sum()
    for p in range(16):
        x = c[:, ph_of == p].ravel()
        print("  phase %2d: %7.0f %7.0f %7.0f %8.0f %8.0f %9.0f   %5.1f %%"
              % (p, x.mean(), *np.percentile(x, [50, 90, 99, 99.9]), x.max(), 100 * x.sum() / tot))
    # units of a 50-frame launch: barrier after a frame unless its phase is 1..7
    cta_time = np.zeros(16); env_time = np.zeros(16); nunits = np.zeros(16)
    total_cta = 0.0; total_env = 0.0
    Gn = G - 1 if heavy else G
    for gi in range(Gn):
        # group gi takes the chunks gi, gi + Gn, ... of 32 consecutive slots after the heavy ones (group_geom)
        T = n - heavy
        chunks = np.arange(gi, (T + 31) // 32, Gn)
        slots = (heavy + chunks[:, None] * 32 + np.arange(32)[None, :]).ravel()
        envs = perm[slots[slots < n]]
        for c0 in range(0, len(envs), 32):
            ce = c[envs[c0:c0 + 32]]                     # [<=32, 200]
            for l0 in range(0, 200, 50):
                f = l0
                while f < l0 + 50:
                    e = f
                    while e < l0 + 49 and 1 <= ph_of[e] <= 7:
                        e += 1
                    u = ce[:, f:e + 1].sum(1)
                    p = ph_of[e]
                    cta_time[p] += u.max() * len(u); env_time[p] += u.sum(); nunits[p] += 1
                    f = e + 1
    total_cta, total_env = cta_time.sum(), env_time.sum()
    print("lockstep replay (non-heavy groups): warp-cycles held by the CTA %.4g, busy %.4g -> busy / held = %.3f"
          % (total_cta, total_env, total_env / total_cta))
    print("  unit ending in phase: share of CTA time, busy/held of those units")
    for p in range(16):
        if nunits[p]:
            print("  phase %2d: %5.1f %%   %.3f" % (p, 100 * cta_time[p] / total_cta, env_time[p] / cta_time[p]))


if __name__ == "__main__":
    main()
