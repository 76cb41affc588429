"""Shared per-operator parity cases (used by tests/test_ops_gpu.py and tools/gpu_op_check.py).

Every distinct conv geometry of Unet-resnet34/50 appears here at a reduced
spatial size / batch so the torch fp32 reference finishes instantly.
"""
# (name, n, h, w, cin, cout, k, stride, pad, relu, residual, in_pitch_extra, out_pitch_extra)
CONV_CASES = [
    ("l1_3x3_64_64",        2, 32, 32,   64,  64, 3, 1, 1, True,  False, 0, 0),
    ("l1_3x3_64_64_res",    2, 32, 32,   64,  64, 3, 1, 1, True,  True,  0, 0),
    ("l2_3x3s2_64_128",     2, 32, 32,   64, 128, 3, 2, 1, True,  False, 0, 0),
    ("l2_ds_1x1s2_64_128",  2, 32, 32,   64, 128, 1, 2, 0, False, False, 0, 0),
    ("l2_3x3_128_128_res",  2, 16, 16,  128, 128, 3, 1, 1, True,  True,  0, 0),
    ("l3_3x3s2_128_256",    2, 16, 16,  128, 256, 3, 2, 1, True,  False, 0, 0),
    ("l3_3x3_256_256_res",  2,  8,  8,  256, 256, 3, 1, 1, True,  True,  0, 0),
    ("l4_3x3s2_256_512",    4,  8,  8,  256, 512, 3, 2, 1, True,  False, 0, 0),
    ("l4_ds_1x1s2_256_512", 4,  8,  8,  256, 512, 1, 2, 0, False, False, 0, 0),
    ("l4_3x3_512_512_res",  4,  4,  4,  512, 512, 3, 1, 1, True,  True,  0, 0),
    ("d0_3x3_768_256",      2,  8,  8,  768, 256, 3, 1, 1, True,  False, 0, 0),
    ("d1_3x3_384_128",      1, 16, 16,  384, 128, 3, 1, 1, True,  False, 0, 0),
    ("d2_3x3_192_64",       1, 32, 32,  192,  64, 3, 1, 1, True,  False, 0, 0),
    ("d3_3x3_128_32",       1, 64, 64,  128,  32, 3, 1, 1, True,  False, 0, 0),
    ("d3_3x3_32_32",        1, 64, 64,   32,  32, 3, 1, 1, True,  False, 0, 0),
    ("d4_3x3_32_16",        1, 64, 128,  32,  16, 3, 1, 1, True,  False, 0, 0),
    ("d4_3x3_16_16",        1, 64, 128,  16,  16, 3, 1, 1, True,  False, 0, 0),
    # pitched views: input is a channel slice of a wider buffer, output lands inside a concat buffer
    ("pitched_in_out",      2, 16, 16,   64,  64, 3, 1, 1, True,  True,  64, 128),
    ("pitched_s2",          2, 16, 16,   64, 128, 3, 2, 1, True,  False, 32, 64),
    # ragged: spatial sizes that do not tile evenly (768-input pyramid: 24, 48, 96)
    ("ragged_24",           3, 24, 24,  128, 128, 3, 1, 1, True,  True,  0, 0),
    ("ragged_24_s2",        3, 24, 24,  128, 256, 3, 2, 1, True,  False, 0, 0),
    ("ragged_48x24",        1, 48, 24,   64,  96, 3, 1, 1, False, False, 0, 0),
    ("tiny_1x1_spatial",    5,  1,  1,  512, 512, 3, 1, 1, True,  True,  0, 0),
    ("tiny_2x2_s2",         3,  2,  2,  256, 512, 3, 2, 1, True,  False, 0, 0),
    # many tiles per persistent CTA (+ resident-weight mode for n_tiles == 1 and small K*N)
    ("persist_16_16",       2, 256, 256,  16,  16, 3, 1, 1, True,  False, 0, 0),
    ("persist_32_16",       1, 256, 384,  32,  16, 3, 1, 1, True,  False, 0, 0),
    ("persist_64_64_res",   4, 128, 128,  64,  64, 3, 1, 1, True,  True,  0, 0),
    ("persist_128_128",     4,  96,  96, 128, 128, 3, 1, 1, True,  True,  0, 0),
    ("persist_64_256_s2",   4, 128, 128,  64, 256, 3, 2, 1, True,  False, 0, 0),
    # resnet50 bottleneck 1x1s and the big decoder K
    ("r50_1x1_64_256",      1, 32, 32,   64, 256, 1, 1, 0, False, False, 0, 0),
    ("r50_1x1_256_64",      1, 32, 32,  256,  64, 1, 1, 0, True,  False, 0, 0),
    ("r50_1x1_1024_2048s2", 1,  8,  8, 1024, 2048, 1, 2, 0, False, False, 0, 0),
    ("r50_1x1_512_2048res", 1,  4,  4,  512, 2048, 1, 1, 0, True,  True,  0, 0),
    ("r50_d0_3x3_3072_256", 1,  8,  8, 3072, 256, 3, 1, 1, True,  False, 0, 0),
]

# Fused nearest-2x upsample + concat + conv3x3 (smp DecoderBlock: interpolate, cat, conv1) — uwm_conv2d_upcat_nhwc_bf16.
# (name, n, h_lo, w_lo, c_x, c_skip, cout, upsample, relu, x_pitch_extra, skip_pitch_extra)
UPCAT_CASES = [
    ("d0_up512_skip256_256", 2,  4,  4, 512, 256, 256, True,  True,  0, 0),
    ("d1_up256_skip128_128", 1,  8,  8, 256, 128, 128, True,  True,  0, 0),
    ("d2_up128_skip64_64",   1, 16, 16, 128,  64,  64, True,  True,  0, 0),
    ("d3_up64_skip64_32",    1, 32, 32,  64,  64,  32, True,  True,  0, 0),
    ("d4_up32_noskip_16",    1, 32, 64,  32,   0,  16, True,  True,  0, 0),
    ("r50_d0_up2048_skip1024", 1, 4, 4, 2048, 1024, 256, True, True, 0, 0),
    ("upcat_ragged_h",       2, 12, 20,  32,  16,  48, True,  False, 0, 0),    # 24x40: H not a multiple of 16, kc=16
    ("upcat_ragged_w",       3,  5,  7,  64,  32,  32, True,  True,  0, 0),    # 10x14: W not a multiple of 8
    ("cat_noup",             1, 16, 24,  64,  64,  64, False, True,  0, 0),
    ("upcat_pitched",        2,  8,  8,  64,  64,  64, True,  True,  32, 64),
    ("upcat_persistent",     2, 128, 128, 32,  0,  16, True,  True,  0, 0),    # many tiles per CTA
]

# Sub-pixel form of conv3x3(nearest_up2x(x)) without a skip — uwm_conv2d_up2x_shuffle_nhwc_bf16.
# (name, n, h_lo, w_lo, cin, cout, relu)
SHUFFLE_CASES = [
    ("d4_up32_16_subpixel",   1, 32, 64, 32, 16, True),
    ("subpixel_ragged",       2, 12, 20, 48, 32, False),     # 24x40 output; 4*32 = 128 GEMM columns
    ("subpixel_64",           1, 16, 16, 64, 64, True),      # 4*64 = 256 GEMM columns (two N tiles possible)
    ("subpixel_persistent",   2, 128, 128, 32, 16, True),
]

# Sub-pixel form of conv3x3(cat(nearest_up2x(x), skip)) - uwm_conv2d_upcat_subpixel_nhwc_bf16.
# (name, n, h_lo, w_lo, c_x, c_skip, cout, relu)
SPX_CASES = [
    ("d3_up64_skip64_32_spx",  1, 32, 32, 64, 64, 32, True),       # the r34 decoder block 3 shape (two sub-tiles)
    ("spx_ragged",             2, 12, 20, 64, 64, 32, False),      # 24x40 output: partial tiles on both axes
    ("spx_narrow_tg1",         1, 16, 8, 64, 128, 16, True),       # one sub-tile per tile, N = 64, two skip chunks
    ("spx_two_x_chunks",       1, 16, 16, 128, 64, 32, True),      # r34 decoder block 2 channels at cout 32
    ("spx_n256",               1, 16, 32, 128, 64, 64, True),      # N = 256 (one sub-tile, 2 x 256 TMEM columns)
    ("spx_persistent",         2, 128, 128, 64, 64, 32, True),     # several tiles per CTA
]

# conv3x3 / head on a 16-channel tensor kept space-to-depth ([n,h,w,4x16] for [n,2h,2w,16]) -
# uwm_conv2d_s2d_nhwc_bf16 / uwm_head_s2d_nhwc_bf16.   (name, n, h_blocks, w_blocks, relu)
S2D_CASES = [
    ("s2d_small",       1, 16, 16, True),
    ("s2d_ragged",      2, 12, 20, False),      # 24x40 pixels: partial tiles on both axes
    ("s2d_wide",        1, 16, 64, True),       # several sub-tiles per tile
    ("s2d_persistent",  2, 128, 128, True),     # 256x256 pixels, several tiles per CTA
]

# Stride-2 conv3x3 (pad 1) on the parity-plane halo kernel - uwm_conv2d_s2_planes_nhwc_bf16.
# (name, n, h_in, w_in, cin, cout, relu)
S2P_CASES = [
    ("l2_0_conv1_s2planes",  1, 32, 32, 64, 128, True),       # r34 layer2.0.conv1 shape
    ("l3_0_conv1_s2planes",  1, 16, 16, 128, 256, True),      # two N tiles
    ("s2planes_ragged",      2, 24, 40, 64, 64, False),       # 12x20 output: partial tiles, bn = 64
    ("s2planes_persistent",  2, 256, 256, 64, 128, True),     # several tiles per CTA
    ("r50_l4_conv2_s2planes", 1, 16, 16, 512, 512, True),     # long K (72 slices), four N tiles
]

# Sub-pixel conv with one N tile per output parity (Cout 128 / 256) and the residual added in the pixel-shuffle epilogue -
# uwm_conv2d_up2x_shuffle_res_nhwc_bf16.   (name, n, h_lo, w_lo, cin, cout, relu, with_residual)
SHUFFLE_RES_CASES = [
    ("d1_x_part_256_128",       1, 16, 16, 256, 128, True, True),     # decoder block 1's upsampled half
    ("d0_x_part_512_256",       2, 8, 8, 512, 256, True, True),       # decoder block 0's upsampled half (N tile = 256)
    ("subpixel_par_ragged",     1, 12, 20, 64, 128, False, False),    # partial tiles, no residual
    ("subpixel_par_persistent", 2, 64, 64, 128, 128, True, True),     # several tiles per CTA
]
