"""Seeded synthetic inputs shared by make_golden.py (reference side, build
container) and the parity tests (oracle / CUDA side).  numpy Generator streams
are platform independent, so the GPU box regenerates bit-identical inputs."""
import numpy as np

# name -> dict(reference argv, feature size, semantic_nc, tail flags)
TAIL_CASES = {
    # BASELINE.json configs[0]: ade20k, 64x64 map, C=256 (match_kernel 1)
    "ade20k_mk1": dict(argv=["--dataset_mode", "ade20k", "--use_attention", "--maskmix", "--PONO", "--PONO_C",
                             "--match_kernel", "1", "--show_warpmask"],
                       img=256, nc=151, kind="uniform", seed=101,
                       flags=dict(match_kernel=1, pono_c=True, warp_mask_losstype="direct")),
    # the reference default match_kernel=3 -> K=2304
    "ade20k_mk3": dict(argv=["--dataset_mode", "ade20k", "--use_attention", "--maskmix", "--PONO", "--PONO_C"],
                       img=256, nc=151, kind="uniform", seed=102,
                       flags=dict(match_kernel=3, pono_c=True)),
    # near one-hot softmax rows (phi = permuted theta + small noise)
    "ade20k_mk1_peaky": dict(argv=["--dataset_mode", "ade20k", "--maskmix", "--PONO", "--PONO_C",
                                   "--match_kernel", "1"],
                             img=256, nc=151, kind="peaky", seed=103,
                             flags=dict(match_kernel=1, pono_c=True)),
    # celebahq-like: spatial centering (no PONO_C), bilinear upsample, cycle
    "celebahq_bilinear_cycle": dict(argv=["--dataset_mode", "celebahq", "--warp_bilinear", "--adaptor_kernel", "4",
                                          "--warp_cycle_w", "1.0", "--match_kernel", "1"],
                                    img=256, nc=19, kind="uniform", seed=104, train=True,
                                    flags=dict(match_kernel=1, pono_c=False, warp_bilinear=True, warp_cycle=True)),
    # deepfashion-like: warp_patch (Cv=48, fold)
    "deepfashion_patch": dict(argv=["--dataset_mode", "deepfashion", "--warp_patch", "--video_like",
                                    "--match_kernel", "1"],
                              img=256, nc=20, kind="uniform", seed=105,
                              flags=dict(match_kernel=1, pono_c=False, warp_patch=True)),
    # small ragged-ish map: 96x96 image -> 24x24 map, N=576 (not a multiple of 128)
    "small_n576_mk3": dict(argv=["--dataset_mode", "ade20k", "--maskmix", "--PONO", "--PONO_C",
                                 "--warp_mask_losstype", "cycle", "--two_cycle", "--warp_cycle_w", "1.0"],
                           img=96, nc=151, kind="uniform", seed=106, train=True,
                           flags=dict(match_kernel=3, pono_c=True, warp_mask_losstype="cycle", warp_cycle=True,
                                      two_cycle=True)),
}


def blocky_onehot(rng, nc, img, block=16, batch=1):
    lab = rng.integers(0, nc, size=(batch, 1, img // block, img // block))
    lab = np.repeat(np.repeat(lab, block, axis=2), block, axis=3)
    oh = np.zeros((batch, nc, img, img), np.float32)
    np.put_along_axis(oh, lab, 1.0, axis=1)
    return oh


def tail_inputs(case, batch=1, c=256):
    """theta_conv, phi_conv [B,256,h,w] fp32; ref_img, real_img [B,3,H,W];
    seg, ref_seg [B,nc,H,W]."""
    spec = TAIL_CASES[case]
    rng = np.random.default_rng(spec["seed"])
    img = spec["img"]
    fh = img // 4
    theta = rng.standard_normal((batch, c, fh, fh)).astype(np.float32)
    if spec["kind"] == "peaky":
        perm = rng.permutation(fh * fh)
        phi = theta.reshape(batch, c, -1)[:, :, perm].reshape(theta.shape)
        phi = (phi + 0.05 * rng.standard_normal(theta.shape)).astype(np.float32)
    else:
        # smooth-ish second field so the softmax is neither flat nor one-hot
        phi = (0.6 * theta[:, :, ::-1, :] + 0.8 * rng.standard_normal(theta.shape)).astype(np.float32)
    ref_img = rng.uniform(-1, 1, (batch, 3, img, img)).astype(np.float32)
    real_img = rng.uniform(-1, 1, (batch, 3, img, img)).astype(np.float32)
    nc = spec["nc"]
    if nc == 20:  # deepfashion: float pose maps, not one-hot
        seg = rng.uniform(0, 1, (batch, nc, img, img)).astype(np.float32)
        ref_seg = rng.uniform(0, 1, (batch, nc, img, img)).astype(np.float32)
    else:
        seg = blocky_onehot(rng, nc, img, block=16 if img % 16 == 0 else 8, batch=batch)
        ref_seg = blocky_onehot(rng, nc, img, block=16 if img % 16 == 0 else 8, batch=batch)
    return dict(theta=theta, phi=phi, ref_img=ref_img, real_img=real_img, seg=seg, ref_seg=ref_seg)
