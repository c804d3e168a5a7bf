"""chainer_vq_vae_b200: the B200-native hot path of dhgrs/chainer-VQ-VAE.

Importing this package loads csrc/libvqw.so and raises if it is missing -- there is no CPU
or eager fallback."""
from . import _lib
from ._lib import MODES, VqwError, launch_count
from .functions import conv, embed_gather, residual_stack, straight_through, vq_lookup
from .links import Convolution2D, DilatedConvolution2D, EmbedID, namedparams
from .generate import generate_utterance, output_to_wave, write_wav
from .losses import logistic_loss, softmax_cross_entropy
from .net import VAE, ConditionEmbed, Encoder
from .updaters import Adam, GradBucket, VQVAE_ParallelUpdater, VQVAE_StandardUpdater
from .snapshot import load_chainer_snapshot, load_optimizer_state, save_chainer_snapshot
from .utils import VQ, ExponentialMovingAverage, MuLaw
from .wavenet import ResidualBlock, ResidualNet, WaveNet

__all__ = [
    "Encoder", "ConditionEmbed", "VAE", "VQ", "straight_through", "ExponentialMovingAverage",
    "MuLaw", "ResidualBlock", "ResidualNet", "WaveNet", "VQVAE_StandardUpdater",
    "VQVAE_ParallelUpdater", "Adam", "GradBucket", "softmax_cross_entropy", "logistic_loss",
    "conv", "embed_gather", "residual_stack", "vq_lookup", "Convolution2D",
    "DilatedConvolution2D", "EmbedID", "namedparams", "MODES", "VqwError", "launch_count",
    "generate_utterance", "output_to_wave", "write_wav", "load_chainer_snapshot", "save_chainer_snapshot", "load_optimizer_state",
]
