"""A small event model of the fused tower's schedule (DESIGN.md "The hand-off between layers"): pairs of CTAs take the units of all layers round-robin, a
unit's K-blocks wait for the rows of the previous layer (three neighbouring row groups) plus a hand-off latency. It reproduces why the wide tower (1.35 units
per pair and layer at config 2) is bound by  layers x (unit time + hand-off)  rather than by its arithmetic, and how much each thousand cycles of hand-off
is worth. Cycle figures are in thousands; T = one wide unit's MMAs (with its barrier overhead), E = last MMA -> rows published, L = first K-block's TMA,
poll = counter round trip.

    python profiles/tower_schedule_model.py

Calibration against the per-CTA counters of profiles/r2_wide_ab.log / r2_epi_ab.log / r2_kbo_ab.log (issuer total, wait for input, per launch):
    old epilogue            E = 12.3: measured 493 k / 190 k
    TMA-store epilogue      E =  6.4: measured 404 k /  98 k
    tile 0 first            E =  4.1: measured 378 k /  71 k
    per-64-channel counters E =  2.7: measured 352 k /  46 k
The model is deterministic (no spread in the producers' finishing times) and so sits 4-9 % below the measured totals; the slope per thousand cycles of E agrees.
With E = 0 it still needs 325 k (poll + TMA of the first K-block remain), against 317 k of pure arithmetic: what is left in the hand-off is worth ~6 %."""


def sim(nc=74, groups=50, nh=2, layers=13, T=18.4, E=4.1, L=3.5, poll=1.0, stem=0.25, ring=3, nkb=4):
    units = groups * nh
    rotate = nc - units % nc  # engine.cu: tower_rotation
    pub = {}
    pair_t = [0.0] * nc
    slot_free = [[0.0] * ring for _ in range(nc)]
    slot_idx = [0] * nc
    wait_tot = 0.0
    for layer in range(layers):
        for q in range((units + nc - 1) // nc):
            for c in range(nc):
                u = (c + layer * rotate) % nc + q * nc
                if u >= units:
                    continue
                g, h = u // nh, u % nh
                kbs = 1 if layer == 0 else nkb  # the stem reads one K-block (18 planes padded to 64)
                tk = T / nkb
                for kb in range(kbs):
                    dep = 0.0
                    if layer > 0:
                        dep = max(pub[(layer - 1, gg, h2)] for gg in (g - 1, g, g + 1) if 0 <= gg < groups for h2 in range(nh)) + poll
                    s = slot_idx[c]
                    ld = max(dep, slot_free[c][s]) + L
                    wait_tot += max(0.0, ld - pair_t[c])
                    end = max(pair_t[c], ld) + tk
                    pair_t[c] = end
                    slot_free[c][s] = end
                    slot_idx[c] = (s + 1) % ring
                pub[(layer, g, h)] = pair_t[c] + E
    return max(pair_t), wait_tot / nc


if __name__ == "__main__":
    print("config 2, wide tower (100 units per layer, 74 pairs, 13 layers); cycles in thousands")
    for E in (12.3, 6.4, 4.1, 2.7, 1.5, 0.0):
        total, wait = sim(E=E)
        print(f"  E = {E:4.1f}: launch {total:5.0f} k, issuer waits for input {wait:4.0f} k")
    total, wait = sim(E=0.0, L=0.0, poll=0.0)
    print(f"  no hand-off at all: launch {total:5.0f} k (= the arithmetic: 17.6 passes x 18.4 k)")
    print("config 4 (200 units per layer, 41 layers): every unit's rows were finished a pass earlier")
    for E in (12.3, 2.7):
        total, wait = sim(groups=100, layers=41, E=E)
        print(f"  E = {E:4.1f}: launch {total:5.0f} k, issuer waits for input {wait:4.0f} k")
