"""One-screen summary of a bench.py JSON line (tuning aid).   python tools/bench_summary.py FILE [FILE...]"""
import json
import sys


def g(d, *ks, default=None):
    for k in ks:
        if not isinstance(d, dict) or k not in d or d[k] is None:
            return default
        d = d[k]
    return d


for path in sys.argv[1:]:
    d = json.loads(open(path).read().strip().splitlines()[-1])
    print(f"== {path}: N={d.get('n_gpus')} steps={d.get('steps')}")
    print(f"value {d['value']:.4g} ({d['ms_per_step']:.3f} ms)  sustained {g(d,'sustained','value',default=0):.4g}  e2e {d['e2e']['value']:.4g} "
          f"({d['e2e']['ms_per_step']:.3f} ms, {g(d,'e2e','frac_of_ceiling',default=0):.2f} of the {g(d,'e2e','h2d_ceiling_gbs',default=0):.1f} GB/s ceiling)  launches {d.get('gpu_launches')}")
    r, p = d['roofline'], d['roofline_pyramid']
    print(f"solver {r['ms_per_launch']:.3f} ms frac {r['frac']:.3f} (peak {r['peak']:.1f} T)  pyramid {p['ms_per_launch']:.3f} ms frac {p['frac']:.3f}")
    print(f"subpixel {g(d,'subpixel','value',default=0):.4g} ({g(d,'subpixel','ms_per_step',default=0):.3f} ms) e2e {g(d,'subpixel','e2e','value',default=0):.4g}   "
          f"8x8 {g(d,'other_patches','8x8','value',default=0):.4g}  11x11 {g(d,'other_patches','11x11','value',default=0):.4g}")
    print(f"C4 fwd {g(d,'config_c4','forward','value',default=0):.4g} inv {g(d,'config_c4','inverse','value',default=0):.4g}   "
          f"C1 {g(d,'single_call','c1_150_features','forward','ms_per_call',default=0):.3f} ms/call  C2 {g(d,'sequence_mode','handles','ms_per_frame',default=0):.3f} ms/frame   "
          f"cpu {g(d,'cpu_baseline','value',default=0):.4g} ({g(d,'cpu_baseline','kind')})")
    for name in ('vs_exact_kernel', 'vs_cpu_arm_full_batch', 'subpixel_vs_cpu_arm_32_pairs'):
        q = g(d, 'parity', name)
        if q:
            print(f"parity {name}: flags {q['flag_mismatches']} max {q['max_abs_dpos_px']:.2e} over {q['n_over_1e-3_px']} bit-identical {q['bit_identical_fraction']:.5f}")
    print(f"guess = truth + N(0,2): {g(d,'guess_projected','value',default=0):.4g} iters {g(d,'guess_projected','gn_iters_per_level')} accuracy {g(d,'guess_projected','accuracy_vs_truth')}")
    print(f"accuracy vs truth (kp2 = kp1): {d.get('accuracy_vs_truth')}   GN passes exact/this {g(d,'parity','vs_exact_kernel','gn_passes_per_level_exact_kernel')} / {g(d,'parity','vs_exact_kernel','gn_passes_per_level_this_kernel')}")
    fc = g(d, 'frontend_chain', default={}) or {}
    print(f"frontend chain: detect {fc.get('ms_detect', 0):.3f} + track {fc.get('ms_track', 0):.3f} + triangulate {fc.get('ms_triangulate', 0):.3f} ms per {d['config'].get('pairs_per_gpu')} frames = {fc.get('frames_per_s', 0):.4g} frames/s, {fc.get('corners_per_frame_mean', 0):.0f} corners/frame")
    pts = g(d, 'sweep_c5', 'points', default=[])
    print("sweep: " + "  ".join(f"{q['features_per_pair']}/{q['patch']}:{q['value']:.3g}" for q in pts))
    print(f"clocks {g(d,'clocks','sm_mhz')} / {g(d,'clocks','sm_max_mhz')} {g(d,'clocks','reasons')}   per-rank resident {g(d,'per_rank','resident_ms_per_step')}")
