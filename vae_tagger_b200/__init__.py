"""vae_tagger_b200 -- B200-native (sm_100a) implementation of vae-tagger's encode+tag hot path.

The package mirrors the reference's Python surface for that path
(``diffusers_vae_loader``, ``modules``, ``improved_losses``, ``infer_full``,
``train_decoder``) and backs it with hand-written CUDA kernels behind a C-ABI
(``include/vae_tagger_b200.h``).  There is no CPU / PyTorch fallback on the product path.
"""
__version__ = "0.1.0"
