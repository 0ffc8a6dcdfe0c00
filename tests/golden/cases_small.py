"""Flag sets for the 64x64 model-level goldens (tests/golden/small_*.npz): one generator + discriminator pass of the
UNMODIFIED reference at batch 1, crop 64, seeded init and inputs.  Cheap enough on both sides (reference: minted once in
the build container by make_golden_small.py; ours: ~10 s per case on CPU) to cover the option branches the three 256x256
goldens do not reach.  Every name -> reference argv (without the size / batch flags CROP adds)."""

CROP = ["--crop_size", "64", "--load_size", "64", "--batchSize", "1"]
_ADE = ["--dataset_mode", "ade20k", "--PONO", "--PONO_C"]

SMALL_CASES = {
    # residual stack of 256 + 151 + 2 = 409 channels (C % 8 == 1)
    "coordconv_maskmix": _ADE + ["--maskmix", "--use_coordconv"],
    # 256 + 19 = 275 channels, batch-statistics SPADE (the SynchronizedBatchNorm stub), bilinear warp, cycle, SAGAN block
    "celebahq_maskmix_attn_cycle": ["--dataset_mode", "celebahq", "--maskmix", "--use_attention", "--warp_bilinear",
                                    "--warp_cycle_w", "0.1"],
    # 256 + 20 = 276 channels, float pose maps
    "deepfashion_maskmix_videolike": ["--dataset_mode", "deepfashion", "--maskmix", "--video_like", "--PONO", "--PONO_C"],
    # column-softmax mask loss (correspondence.py:337-346, pix2pix_model.py:261-276)
    "cycle_mask": _ADE + ["--maskmix", "--warp_mask_losstype", "cycle"],
    # 4x4 adaptor kernels, edge maps as labels, both cycle terms (correspondence.py:350-372)
    "celebahqedge_two_cycle": ["--dataset_mode", "celebahqedge", "--PONO", "--PONO_C", "--adaptor_kernel", "4",
                               "--warp_cycle_w", "1.0", "--two_cycle"],
    "match_kernel_1": _ADE + ["--match_kernel", "1"],
    "warp_stride_2": _ADE + ["--warp_stride", "2"],
    "perceptual_4_2_ctx22": _ADE + ["--which_perceptual", "4_2", "--weight_perceptual", "0.001", "--use_22ctx"],
    "lsgan_no_feat": _ADE + ["--gan_mode", "ls", "--no_ganFeat_loss"],
    "original_gan_fm_ratio": _ADE + ["--gan_mode", "original", "--fm_ratio", "0.5", "--warp_self_w", "100.0"],
    "adaptor_se_deeper": _ADE + ["--adaptor_se", "--adaptor_res_deeper"],
    "adaptor_nonlocal_dilation": _ADE + ["--adaptor_nonlocal", "--dilation_conv"],
    "eqlr_sn": _ADE + ["--eqlr_sn"],
    "d_cam": _ADE + ["--D_cam", "1.0"],
    "domain_classifier": _ADE + ["--weight_domainC", "1.0"],
    "cbn_mask": _ADE + ["--CBN_intype", "mask", "--maskmix", "--use_attention"],
    # instance statistics in SPADE (at batch 1 the same numbers as the default batch statistics: another kernel path)
    "no_pono_instance_stats": ["--dataset_mode", "ade20k", "--norm_G", "spectralspadeinstance3x3"],
}

# the module-only variants (the tape does not take these layers: they pin the mirror modules, not the kernels' host
# side) run when COCOS_ALL_SMALL_CASES=1; the default CPU suite keeps to the cases that reach the tape
EXTENDED_ONLY = ("original_gan_fm_ratio", "adaptor_se_deeper", "adaptor_nonlocal_dilation", "eqlr_sn", "d_cam",
                 "domain_classifier")
