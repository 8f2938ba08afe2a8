"""The caller of the hot path: a small pre-norm ViT with injected attention / RPE plugins.

Behaviourally equal to the reference's models/core/base_vit.py and models/components/unified_transformer.py
(same module names -> same state_dict keys, same init laws), so reference checkpoints load.  At the reference's
dims (dim 32, MLP 64) the block around the attention core, the patch embedding (4x4 patches) and the head + loss run as
fused kernels of this library (`_forward_fused`, `_fused_ends`); other geometries fall back to the op-by-op modules
(Linear / LayerNorm / GELU on cuBLAS / ATen) around `block.attention`, which is always ours.
"""
from typing import Callable, Dict, Optional

import torch
import torch.nn as nn

from . import ops
from .nn import LayerNorm, Linear


def _autocast_ok() -> bool:
    """No autocast, or bf16 autocast on CUDA (the fused fp32 kernels then feed the attention core bf16 operands)."""
    return (not torch.is_autocast_enabled()) or torch.get_autocast_dtype("cuda") == torch.bfloat16


class UnifiedTransformerBlock(nn.Module):
    """x + attn(norm1(x), rpe) ; x + mlp(norm2(x))   (unified_transformer.py:64-90)."""

    def __init__(self, dim: int, attention: nn.Module, rpe: Optional[nn.Module] = None, mlp_dim: int = None,
                 dropout: float = 0.0):
        super().__init__()
        self.dim = dim
        self.mlp_dim = mlp_dim or dim * 4
        self.attention = attention
        self.rpe = rpe
        self.mlp = nn.Sequential(Linear(dim, self.mlp_dim), nn.GELU(), nn.Dropout(dropout),
                                 Linear(self.mlp_dim, dim), nn.Dropout(dropout))
        self.norm1 = LayerNorm(dim)
        self.norm2 = LayerNorm(dim)

    def forward(self, x: torch.Tensor, drop_seed=None, drop_salt: int = 0) -> torch.Tensor:
        """drop_seed / drop_salt (optional, fused path only): a device-resident dropout seed shared by the caller's blocks and
        this block's index, so that a model draws ONE seed per forward pass instead of one per block (BaseViT.features)."""
        if self._fused(x):
            return self._forward_fused(x, drop_seed, drop_salt)
        x = x + self.attention(self.norm1(x), rpe=self.rpe)  # the RPE goes INTO the attention
        return x + self.mlp(self.norm2(x))

    # ---- fused path (csrc/erv_block_fused.cu): two kernels around the attention core instead of ~14 library ops
    def _fused(self, x: torch.Tensor) -> bool:
        att = self.attention
        if not (ops.FUSED_BLOCK and x.is_cuda and x.dtype == torch.float32 and _autocast_ok()):
            return False
        if not (hasattr(att, "core") and hasattr(att, "qkv") and hasattr(att, "proj") and hasattr(att, "proj_dropout")):
            return False
        if att.qkv.weight.dtype != torch.float32 or att.proj.bias is None or not ops.block_supported(self.dim, self.mlp_dim):
            return False
        # p >= 1 (nn.Dropout accepts 1.0: everything dropped) has no finite 1/(1-p): leave it to the op-by-op modules
        return att.proj_dropout.p == self.mlp[2].p == self.mlp[4].p and att.proj_dropout.p < 1.0

    def _forward_fused(self, x: torch.Tensor, drop_seed=None, drop_salt: int = 0) -> torch.Tensor:
        att = self.attention
        att.before_qkv(x.shape, self.rpe)
        # bf16 autocast: the attention core sees the bf16 qkv an autocast Linear would hand it (written directly by the
        # kernel); LayerNorm, the projections and the MLP stay in fp32 (more accurate than the reference's autocast run)
        qkv, x = ops.block_ln_qkv(x, self.norm1.weight, self.norm1.bias, att.qkv.weight, att.qkv.bias, self.norm1.eps,
                                  with_residual=True, out_dtype=torch.bfloat16 if torch.is_autocast_enabled() else None)
        a = att.core(qkv, x.shape, self.rpe)
        p = self.mlp[2].p if self.training else 0.0
        seed = (drop_seed if drop_seed is not None else ops.dropout_seed(x.device)) if p > 0 else None
        return ops.block_mlp(a, x, att.proj.weight, att.proj.bias, self.norm2.weight, self.norm2.bias, self.mlp[0].weight,
                             self.mlp[0].bias, self.mlp[3].weight, self.mlp[3].bias, self.norm2.eps, p, seed, drop_salt)

    def extra_repr(self) -> str:
        return f"dim={self.dim}, mlp_dim={self.mlp_dim}, has_rpe={self.rpe is not None}"


class BaseViT(nn.Module):
    def __init__(self, image_size: int, in_channels: int, patch_size: int, num_classes: int, dim: int, depth: int,
                 heads: int, mlp_dim: int, dropout: float = 0.1, attention_builder: Optional[Callable] = None,
                 rpe_builder: Optional[Callable] = None):
        super().__init__()
        assert image_size % patch_size == 0, f"Image size {image_size} must be divisible by patch size {patch_size}"
        assert dim % heads == 0, f"Model dimension {dim} must be divisible by number of heads {heads}"
        if attention_builder is None:
            raise ValueError("attention_builder must be provided")
        self.image_size, self.patch_size, self.in_channels = image_size, patch_size, in_channels
        self.num_classes, self.dim, self.depth, self.heads = num_classes, dim, depth, heads
        self.mlp_dim, self.dropout = mlp_dim, dropout
        self.num_patches = (image_size // patch_size) ** 2
        self.patch_dim = in_channels * patch_size * patch_size

        self.patch_embedding = Linear(self.patch_dim, dim)
        self.cls_token = nn.Parameter(torch.randn(1, 1, dim))
        self.pos_embedding = nn.Parameter(torch.randn(1, self.num_patches + 1, dim))
        blocks = []
        for _ in range(depth):
            attention = attention_builder(dim=dim, heads=heads, dropout=dropout)
            # one RPE instance per block; the sequence has num_patches + 1 tokens (CLS)   base_vit.py:138-142
            rpe = rpe_builder(num_patches=self.num_patches + 1, dim=dim, heads=heads) if rpe_builder else None
            blocks.append(UnifiedTransformerBlock(dim, attention, rpe, mlp_dim, dropout))
        self.transformer_blocks = nn.ModuleList(blocks)
        self.mlp_head = nn.Sequential(LayerNorm(dim), Linear(dim, num_classes))
        self._init_weights()

    def _init_weights(self):  # base_vit.py:153-172
        nn.init.normal_(self.pos_embedding, std=0.02)
        nn.init.normal_(self.cls_token, std=0.02)
        for m in self.modules():
            if isinstance(m, nn.Linear):
                nn.init.xavier_uniform_(m.weight)
                if m.bias is not None:
                    nn.init.constant_(m.bias, 0)
            elif isinstance(m, nn.LayerNorm):
                nn.init.constant_(m.bias, 0)
                nn.init.constant_(m.weight, 1.0)

    def patchify(self, x: torch.Tensor) -> torch.Tensor:
        """[B, C, H, W] -> [B, patches, C*p*p], patches in raster order (base_vit.py:174-198)."""
        b, c, h, w = x.shape
        assert c == self.in_channels, f"Expected {self.in_channels} channels, got {c}"
        assert h == self.image_size and w == self.image_size, \
            f"Expected {self.image_size}x{self.image_size} images, got {h}x{w}"
        p = self.patch_size
        x = x.reshape(b, c, h // p, p, w // p, p).permute(0, 2, 4, 1, 3, 5)
        return x.reshape(b, self.num_patches, self.patch_dim)

    def _fused_ends(self, x: torch.Tensor) -> bool:
        """The embedding / head+loss kernels of csrc/erv_embed_head.cu apply (fp32 CUDA, dim 32, no autocast)."""
        return (ops.FUSED_BLOCK and x.is_cuda and x.dtype == torch.float32 and _autocast_ok()
                and self.patch_embedding.weight.dtype == torch.float32 and self.patch_embedding.bias is not None
                and ops.embed_supported(self.dim, self.patch_dim))

    def features(self, x: torch.Tensor) -> torch.Tensor:
        """images -> token features after the last block [B, N, dim]."""
        if (ops.FUSED_EMBED and self._fused_ends(x)
                and (ops.embed_fast(self.in_channels, self.patch_size, self.num_patches + 1) or ops.FUSED_EMBED == "always")):
            b, c, h, w = x.shape
            assert c == self.in_channels, f"Expected {self.in_channels} channels, got {c}"
            assert h == self.image_size and w == self.image_size, \
                f"Expected {self.image_size}x{self.image_size} images, got {h}x{w}"
            x = ops.embed(x, self.patch_embedding.weight, self.patch_embedding.bias, self.cls_token, self.pos_embedding,
                          self.patch_size)
        else:
            x = self.patch_embedding(self.patchify(x))
            x = torch.cat([self.cls_token.expand(x.shape[0], -1, -1), x], dim=1) + self.pos_embedding
        cut = getattr(self, "_cut_after", None)
        # one device-resident dropout seed per forward pass; block i salts it with its index (the mask hash takes
        # (seed, salt, element)), instead of a clone + add kernel pair per block
        seed = None
        if self.training and x.is_cuda and any(getattr(b, "mlp", None) is not None and b.mlp[2].p > 0
                                                for b in self.transformer_blocks):
            seed = ops.dropout_seed(x.device)
        for i, block in enumerate(self.transformer_blocks):
            x = block(x, seed, 0x1000 + i) if seed is not None else block(x)
            if cut == i and x.requires_grad:  # erv_b200.train: the backward is split here to overlap the gradient all-reduce
                self._cut_tensor = x
        return x

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return self.mlp_head(self.features(x)[:, 0])

    def loss(self, images: torch.Tensor, labels: torch.Tensor) -> torch.Tensor:
        """Mean cross-entropy of forward(images) against labels (the reference's training criterion,
        experiments/utils/training.py:57-60), with the final LayerNorm, classifier and loss in one kernel when possible."""
        x = self.features(images)
        ln, fc = self.mlp_head[0], self.mlp_head[1]
        if self._fused_ends(images) and fc.bias is not None and fc.out_features <= 32:
            return ops.head_loss(x, ln.weight, ln.bias, fc.weight, fc.bias, labels, ln.eps)
        return torch.nn.functional.cross_entropy(self.mlp_head(x[:, 0]).float(), labels)

    def count_parameters(self) -> Dict[str, int]:
        total = sum(p.numel() for p in self.parameters())
        trainable = sum(p.numel() for p in self.parameters() if p.requires_grad)
        return {"total": total, "trainable": trainable, "non_trainable": total - trainable}

    def get_attention_maps(self):
        raise NotImplementedError("Attention map extraction must be implemented by specific model variants")

    def extra_repr(self) -> str:
        return (f"image_size={self.image_size}, patch_size={self.patch_size}, num_patches={self.num_patches}, "
                f"dim={self.dim}, depth={self.depth}, heads={self.heads}")
