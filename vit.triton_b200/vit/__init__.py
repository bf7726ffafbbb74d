"""B200-native ViT forward behind the host API of cmeraki/vit.triton (``vit.vit.VIT``,
``vit.kernels.*``, ``vit.utils.transfer_pretrained_weights``)."""
