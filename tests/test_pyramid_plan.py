"""CPU tests of the shipped host-side plan of the fused DWT pyramid kernel (pytorch_wavelets_b200/csrc/pyramid_plan.h,
compiled with g++ into the test emulation library): shared-memory layout invariants and the ring-depth rule that the
kernel's dataflow relies on (every group a consumer stage needs is resident together with room for the producer)."""
import itertools

import pytest

from tests.emu import emu_backend as eb

MAX_SMEM = 227 * 1024


def _hs(L):
    m = 1
    while (L // 2) * m < 4 or ((L // 2) * m) & 1:
        m += 1
    return (L // 2) * m


@pytest.mark.parametrize('L', [2, 4, 6, 8, 10, 12])
@pytest.mark.parametrize('mode', ['zero', 'symmetric', 'reflect'])
def test_plan_layout_invariants(L, mode):
    for (H, W), J in itertools.product([(64, 64), (96, 128), (512, 512), (200, 256), (1024, 1024), (37, 52)], [1, 2, 3, 4]):
        d = eb.plan_pyramid(16, H, W, J, L, mode)
        if d is None:
            continue
        assert d['smem_bytes'] <= MAX_SMEM and d['threads'] <= 512 and d['threads'] % 32 == 0
        regions = []
        h, w = H, W
        warp = 2
        for l, v in enumerate(d['levels']):
            assert (v['H'], v['W']) == (h, w)
            assert v['Ho'] == (h + L - 1) // 2 and v['Wo'] == (w + L - 1) // 2
            assert v['warp0'] == warp and v['nwarps'] == -(-(-(-v['Wo'] // 3)) // 32)
            warp += v['nwarps']
            halo = (L - 2 + 3) // 4 * 4
            # a lane that owns a valid column reads 2*3 + L - 2 floats starting at halo + 2*c0 - (L-2)
            assert v['in_pitch'] >= halo + 2 * (v['Wo'] - 1) + 2 * 3 and v['in_pitch'] >= halo + w + L - 1
            assert v['in_pitch'] % 4 == 0 and v['in_off'] % 4 == 0 and v['st_off'] % 4 == 0 and v['st_cap'] % 4 == 0
            regions.append((v['in_off'], v['in_off'] + v['in_rows'] * v['in_pitch']))
            nb = 4 if l == J - 1 else 3
            assert v['nbands'] == nb
            cap_ll = 2 * _hs(L) // d['split'] * d['ll_pitch'] if l == J - 1 else 0
            regions.append((v['st_off'], v['st_off'] + 3 * v['st_cap'] + cap_ll))
            h, w = v['Ho'], v['Wo']
        assert d['threads'] == 32 * warp
        regions.append((d['zero_off'], d['zero_off'] + max(v['in_pitch'] for v in d['levels'])))
        regions.sort()
        assert regions[0][0] * 4 >= 8 * d['n_bars']                      # barriers sit in front of everything
        for (a0, a1), (b0, b1) in zip(regions, regions[1:]):
            assert a1 <= b0, 'shared-memory regions overlap'
        assert regions[-1][1] * 4 <= d['smem_bytes']


@pytest.mark.parametrize('L', [4, 8, 12])
@pytest.mark.parametrize('mode', ['zero', 'symmetric', 'reflect'])
def test_ring_depth_covers_every_consumer_stage(L, mode):
    """Replay the consumer's stage sequence (the same rules as pyr_stage_max_row / pyr_stage_release_bound): the groups
    a stage needs, counted from the oldest group not yet released, must fit the ring with one group to spare."""
    HS, PL, PRO = _hs(L), L - 2, (L - 2) // 2
    RS = 2 * HS

    def ext(i, n):
        if 0 <= i < n:
            return i
        if mode == 'zero':
            return -1
        if mode == 'symmetric':
            r = i % (2 * n)
            return r if r < n else 2 * n - 1 - r
        r = i % (2 * n - 2)
        return r if r < n else 2 * n - 2 - r

    for H in (64, 99, 130, 512):
        d = eb.plan_pyramid(4, H, 256, 3, L, mode)
        if d is None:
            continue
        for l in range(1, 3):
            u, v = d['levels'][l - 1], d['levels'][l]
            h = v['H']
            rel = 0
            for t in range(v['n_stage']):
                rows = [ext(e, h) for e in range(t * RS - PL, (t + 1) * RS - PL)]
                need = min((max([r for r in rows if r >= 0] + [0]) + PRO) // HS, u['n_stage'] - 1)
                assert need - rel + 1 <= v['n_in'] - 1, (H, l, t)
                lo = min((t + 1) * RS - PL, h - L + 1)
                # rows still needed by later stages must not be released
                # (extended rows beyond 2*Ho - 1 only feed half-stages that emit nothing: what they read is never used)
                later = [ext(e, h) for e in range((t + 1) * RS - PL, min(v['n_stage'] * RS - PL, 2 * v['Ho']))]
                later = [r for r in later if r >= 0]
                while rel < u['n_stage'] and (rel + 1) * HS - PRO <= lo:
                    rel += 1
                if later:
                    assert min(later) >= min(rel * HS - PRO, h), 'released a row a later stage reads'


def test_policy_shapes_of_the_baseline_configuration():
    d = eb.plan_pyramid(4096, 512, 512, 1, 8, 'symmetric', ll_pitch=288)     # level 1 of BASELINE configs[1]
    assert d is not None and d['threads'] == 160 and d['ll_pitch'] == 288
    assert (233472 // (d['smem_bytes'] + 1024)) >= 3                           # three CTAs per SM
    assert eb.plan_pyramid(256, 2048, 2048, 4, 16, 'zero') is None             # configs[4] does not fit: level kernels
    assert eb.plan_pyramid(16, 512, 510, 1, 8, 'symmetric') is None            # rows must be 16-byte multiples (TMA)
    assert eb.plan_pyramid(16, 512, 512, 1, 8, 'periodization') is None
