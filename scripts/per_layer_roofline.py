"""Per-launch roofline of one bench step: the ncu launch list (profiles/r2_launches_bench_c2.csv, last 160 launches) joined with
the analytic FLOPs and minimum HBM bytes of every launch of BigGAN-deep-256 forward + alex-LPIPS + backward for 18 candidates
(the launch order of pix2latent_b200/csrc/biggan.cu / lpips.cu).

    python scripts/per_layer_roofline.py profiles/r2_launches_bench_c2.csv profiles/r2_per_layer_roofline.md

roofline time of a launch = max(FLOPs / tensor peak, bytes / HBM peak) with the MEASURED peaks (MEASURED_PEAKS.json, else
the values below); `frac` = roofline time / measured duration. Durations under ncu are serialised and cold-cache."""
import csv
import json
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
B = 18
E = 2  # bytes per 16-bit element


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d.get("bf16_tflops_sustained", 1381.3), d.get("hbm_gbs", 6438.5)
    return 1381.3, 6438.5


def sequence():
    """[(label, kind, flops, bytes)] in launch order; kind 'conv' | 'glue'"""
    ch = 128
    layers = [(0, 16, 16), (1, 16, 16), (0, 16, 16), (1, 16, 8), (0, 8, 8), (1, 8, 8), (0, 8, 8), (1, 8, 4), (0, 4, 4), (1, 4, 2), (0, 2, 2),
              (1, 2, 1)]
    blocks, H = [], 4
    for up, ci, co in layers:
        blocks.append(dict(up=up, cin=ch * ci, cout=ch * co, mid=ch * ci // 4, Hin=H, Hout=2 * H if up else H))
        H = 2 * H if up else H
    seq = []

    def conv(label, M, N, K, rd, wr):
        seq.append((label, "conv", 2.0 * M * N * K, float(rd + wr + K * N * E)))

    def glue(label, nbytes):
        seq.append((label, "glue", 0.0, float(nbytes)))

    def fwd_block(i, bl, last):
        pi, po = B * bl["Hin"] ** 2, B * bl["Hout"] ** 2
        cin, mid, cout = bl["cin"], bl["mid"], bl["cout"]
        conv("blk%d conv_0 1x1" % i, pi, mid, cin, pi * cin * E, (po if bl["up"] else pi) * mid * E + (pi * mid * E if bl["up"] else 0))
        conv("blk%d conv_1 3x3" % i, po, mid, 9 * mid, po * mid * E, po * mid * E)
        conv("blk%d conv_2 3x3" % i, po, mid, 9 * mid, po * mid * E, po * mid * E)
        conv("blk%d conv_3 1x1 +skip" % i, po, cout, mid, po * mid * E + pi * cout * E, po * cout * E * (1 if last else 2))

    def bwd_block(i, bl):
        pi, po = B * bl["Hin"] ** 2, B * bl["Hout"] ** 2
        cin, mid, cout = bl["cin"], bl["mid"], bl["cout"]
        conv("blk%d dgrad conv_3" % i, po, mid, cout, po * cout * E + po * mid * E, po * mid * E)
        conv("blk%d dgrad conv_2" % i, po, mid, 9 * mid, 2 * po * mid * E, po * mid * E)
        conv("blk%d dgrad conv_1" % i, po, mid, 9 * mid, po * mid * E + (0 if bl["up"] else po * mid * E), po * mid * E)
        if bl["up"]:
            glue("blk%d pool+bn/relu bwd" % i, po * mid * E + 2 * pi * mid * E)
            glue("blk%d skip-gradient 2x2 sum" % i, po * cout * E + pi * cout * E)
        conv("blk%d dgrad conv_0 +skip" % i, pi, cin, mid, pi * mid * E + pi * cin * E + pi * cout * E, pi * cin * E)

    glue("concat z,c", 0)
    glue("cond -> BN affine GEMV", 2 * 24192 * 256 * 4)
    glue("final-BN affine", 0)
    glue("gen_z GEMV", 32768 * 256 * 4)
    for i in range(8):
        fwd_block(i, blocks[i], False)
    C, Hh, dq, dv = 512, 64, 64, 256
    px, Nk, nq = B * Hh * Hh, Hh * Hh // 4, 2 * 64 + 256
    conv("attn qkv 1x1 (+ theta^T)", px, nq, C, px * C * E, px * nq * E + px * dq * E)
    glue("attn maxpool phi", px * dq * E * 1.5)
    glue("attn maxpool g", px * dv * E * 1.5)
    conv("attn softmax pass 1: S = theta phi^T -> (max, sum) per row and N tile", px, Nk, dq, px * dq * E + B * Nk * dq * E, px * (Nk // 128) * 8)
    conv("attn softmax pass 2: P = exp(S - M) / L, + P^T", px, Nk, dq, px * dq * E + B * Nk * dq * E, 2 * px * Nk * E)
    conv("attn O = P g", px, dv, Nk, px * Nk * E + B * Nk * dv * E, px * dv * E)
    conv("attn out 1x1 +x", px, C, dv, px * dv * E + px * C * E, 2 * px * C * E)
    for i in range(8, 12):
        fwd_block(i, blocks[i], i == 11)
    R = 256
    conv("rgb head (N=27 tap-expanded)", B * R * R, 27, 128, B * R * R * 128 * E, B * R * R * 27 * 4)
    glue("rgb gather + tanh", B * R * R * (27 + 3) * 4)
    glue("L1 loss", B * R * R * 3 * 4 * 2)
    alex = [(3 * 121, 64, 63, 1, "conv1 (im2col 11x11)"), (64, 192, 31, 25, "conv2 5x5"), (192, 384, 15, 9, "conv3 3x3"),
            (384, 256, 15, 9, "conv4 3x3"), (256, 256, 15, 9, "conv5 3x3")]
    glue("alex im2col", B * R * R * 3 * 4 + B * 63 * 63 * 384 * E)
    for j, (ci, co, h, taps, name) in enumerate(alex):
        K = 384 if j == 0 else ci * taps
        conv("lpips " + name, B * h * h, co, K, B * h * h * (384 if j == 0 else ci) * E, B * h * h * co * E)
        if j in (0, 1):
            glue("lpips maxpool", B * h * h * co * E * 1.3)
    for j in range(5):
        h, co = alex[j][2], alex[j][1]
        glue("lpips distance %d" % j, B * h * h * co * E * 2)
    glue("loss slots -> loss", 0)
    for j in (4, 3, 2, 1, 0):
        ci, co, h, taps, name = alex[j]
        hin = {0: 63, 1: 31, 2: 15, 3: 15, 4: 15}[j]
        K = co * taps if j else 64
        N = 384 if j == 0 else ci
        conv("lpips dgrad " + name, B * hin * hin, N, K, B * h * h * co * E, B * hin * hin * N * E * 2)
        if j in (2, 1):
            glue("lpips maxpool bwd", B * hin * hin * ci * E * 4)
    glue("alex col2im", B * 63 * 63 * 384 * E + B * R * R * 3 * 4)
    glue("rgb gradient im2col", B * R * R * (6 * 4 + 64 * E))
    conv("rgb head dgrad", B * R * R, 128, 64, B * R * R * (64 + 128) * E, B * R * R * 128 * E)
    for i in (11, 10, 9, 8):
        bwd_block(i, blocks[i])
    conv("attn dO = dh Wo (+ dO^T)", px, dv, C, px * C * E, 2 * px * dv * E)
    glue("attn rowsum(dO o O)", 2 * px * dv * E)
    conv("attn dS = P o (dO g^T - rowsum), + dS^T", px, Nk, dv, px * dv * E + B * Nk * dv * E + px * Nk * E, 2 * px * Nk * E)
    conv("attn dtheta = dS phi", px, dq, Nk, px * Nk * E, px * dq * E)
    conv("attn dphi = dS^T theta", B * Nk, dq, Hh * Hh, px * Nk * E, B * Nk * dq * E)
    conv("attn dg = P^T dO", B * Nk, dv, Hh * Hh, px * Nk * E, B * Nk * dv * E)
    glue("attn maxpool bwd phi", px * dq * E * 1.5)
    glue("attn maxpool bwd g", px * dv * E * 1.5)
    conv("attn dx = dqkv Wqkv + dh", px, C, nq, px * nq * E + px * C * E, px * C * E)
    for i in range(7, -1, -1):
        bwd_block(i, blocks[i])
    glue("BN-gradient slots -> sums", 0)
    glue("BN-gradient finalise", 0)
    glue("dcond GEMV (BN tables)", 2 * 24192 * 256 * 4)
    glue("dcond GEMV (gen_z)", 32768 * 256 * 4)
    glue("dcond reduce + split", 0)
    return seq


def main():
    src, out = sys.argv[1], sys.argv[2]
    rows = []
    for r in csv.reader(l for l in open(src) if not l.startswith("==")):
        if len(r) > 10 and r[0].isdigit():
            rows.append((re.sub(r"\(.*", "", r[4]).replace("p2l::", "").replace("void ", "").strip(), float(r[-1]) / 1e3))
    seq = sequence()
    step = rows[-len(seq):]
    tf_peak, bw_peak = peaks()
    lines, tot, tot_roof, cls = [], 0.0, 0.0, {}
    for i, ((name, us), (label, kind, fl, by)) in enumerate(zip(step, seq)):
        is_conv = "conv_gemm" in name or "conv3x3" in name
        assert is_conv == (kind == "conv"), "launch %d: %s does not match %s" % (i, name, label)
        t_fl, t_by = fl / (tf_peak * 1e12) * 1e6, by / (bw_peak * 1e9) * 1e6
        roof = max(t_fl, t_by)
        bound = "tensor" if t_fl >= t_by else "hbm"
        tot += us
        tot_roof += roof
        c = cls.setdefault(kind + ":" + bound, [0, 0.0, 0.0])
        c[0] += 1; c[1] += us; c[2] += roof
        lines.append("| %d | %s | `%s` | %.2f | %.0f | %.1f | %.0f | %.0f | %s | %.1f | %.0f %% |" % (
            i, label, name.replace("_kernel", ""), fl / 1e9, by / 1e6, us, fl / us / 1e6 if us else 0, by / us / 1e3 if us else 0, bound,
            roof, 100 * roof / us if us else 0))
    with open(out, "w") as f:
        f.write("# Per-launch roofline of one bench step (BigGAN-deep-256, 18 candidates, fwd + alex-LPIPS + bwd)\n\n")
        f.write("Source: `%s` (ncu `gpu__time_duration.sum`, serialised, cold cache), joined with analytic FLOPs and minimum HBM bytes per launch "
                "(`scripts/per_layer_roofline.py`). Peaks: %.1f TFLOP/s sustained 16-bit tensor, %.1f GB/s HBM (MEASURED_PEAKS.json). "
                "`roof` = max(FLOPs/peak, bytes/peak); `frac` = roof / measured.\n\n" % (src, tf_peak, bw_peak))
        f.write("Step: **%.3f ms measured, %.3f ms sum of per-launch rooflines (%.0f %%)**\n\n" % (tot / 1e3, tot_roof / 1e3, 100 * tot_roof / tot))
        f.write("| class | launches | measured (us) | roofline (us) | frac |\n|---|---:|---:|---:|---:|\n")
        for k, (n, us, roof) in sorted(cls.items(), key=lambda kv: -kv[1][1]):
            f.write("| %s | %d | %.0f | %.0f | %.0f %% |\n" % (k, n, us, roof, 100 * roof / us if us else 0))
        gaps = sorted(((us - max(fl / (tf_peak * 1e12), by / (bw_peak * 1e9)) * 1e6, i, label, us) for i, ((name, us), (label, kind, fl, by))
                       in enumerate(zip(step, seq))), reverse=True)
        f.write("\nLargest absolute gaps (measured - roofline), i.e. where the next microseconds are:\n\n| # | launch | measured us | gap us |\n|---:|---|---:|---:|\n")
        for g, i, label, us in gaps[:24]:
            f.write("| %d | %s | %.1f | %.1f |\n" % (i, label, us, g))
        f.write("\n| # | launch | kernel | GFLOP | MB | us | TFLOP/s | GB/s | bound | roof us | frac |\n|---:|---|---|---:|---:|---:|---:|---:|---|---:|---:|\n")
        f.write("\n".join(lines) + "\n")
    print(open(out).read()[:3000])


if __name__ == "__main__":
    main()
