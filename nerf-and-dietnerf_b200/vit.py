"""Image embedder of DietNeRF's semantic-consistency loss: a ViT-B/32 feature extractor.

The reference loads ``https://tfhub.dev/sayakpaul/vit_b32_fe/1`` (src/DietNeRF.py:14,75-78) -- a frozen ViT-B/32 that
maps a (224,224,3) image in [-1,1] to a 768-d feature.  There is no network here, so the weights are random-initialised
with that architecture (BASELINE.json configs[4] asks for exactly that): 32x32 patch embedding, class token, learned
position embedding, 12 pre-LayerNorm blocks (12 heads, MLP 3072, GELU), final LayerNorm, class-token feature.

This is a SURVEY 8f-4 "next" row around the hot path, not part of it: plain PyTorch modules (library GEMMs/attention),
fp32 like the reference (which switches the Keras policy to float32 while it builds the embedder), frozen parameters,
differentiable w.r.t. the input image only.  The rendered image it consumes and the gradient it returns go through the
hand-written render / render-backward kernels.
"""
import torch
import torch.nn as nn
import torch.nn.functional as F

EMBEDDER_INPUT_SIZE = 224      # src/DietNeRF.py:15


class _Block(nn.Module):
    def __init__(self, width, heads, mlp_dim):
        super().__init__()
        self.heads = heads
        self.ln1 = nn.LayerNorm(width, eps=1e-6)
        self.qkv = nn.Linear(width, 3 * width)
        self.proj = nn.Linear(width, width)
        self.ln2 = nn.LayerNorm(width, eps=1e-6)
        self.fc1 = nn.Linear(width, mlp_dim)
        self.fc2 = nn.Linear(mlp_dim, width)

    def forward(self, x):
        b, t, c = x.shape
        q, k, v = self.qkv(self.ln1(x)).reshape(b, t, 3, self.heads, c // self.heads).permute(2, 0, 3, 1, 4)
        a = F.scaled_dot_product_attention(q, k, v).transpose(1, 2).reshape(b, t, c)
        x = x + self.proj(a)
        return x + self.fc2(F.gelu(self.fc1(self.ln2(x))))


class ViTB32(nn.Module):
    """ViT-B/32 feature extractor: (B,3,224,224) in [-1,1] -> (B,768)."""

    def __init__(self, image_size=EMBEDDER_INPUT_SIZE, patch=32, width=768, layers=12, heads=12, mlp_dim=3072, seed=0):
        super().__init__()
        gen_state = torch.random.get_rng_state()
        torch.manual_seed(seed)
        n_tokens = (image_size // patch) ** 2 + 1
        self.patch = patch
        # the stride-32 32x32 convolution written as a matmul over flattened patches: stays in fp32 like the reference's
        # float32 embedder (cuDNN convolutions default to TF32)
        self.patch_embed = nn.Linear(3 * patch * patch, width)
        self.cls = nn.Parameter(torch.zeros(1, 1, width))
        self.pos = nn.Parameter(torch.randn(1, n_tokens, width) * 0.02)
        self.blocks = nn.ModuleList(_Block(width, heads, mlp_dim) for _ in range(layers))
        self.ln = nn.LayerNorm(width, eps=1e-6)
        torch.random.set_rng_state(gen_state)
        for p in self.parameters():            # hub.KerasLayer(..., trainable=False)
            p.requires_grad_(False)

    def forward(self, images):
        b, c, h, w = images.shape
        p = self.patch
        x = images.reshape(b, c, h // p, p, w // p, p).permute(0, 2, 4, 1, 3, 5).reshape(b, (h // p) * (w // p), c * p * p)
        x = self.patch_embed(x)
        x = torch.cat([self.cls.expand(x.shape[0], -1, -1), x], dim=1) + self.pos
        for blk in self.blocks:
            x = blk(x)
        return self.ln(x)[:, 0]


def embedder_preprocess(images):
    """src/DietNeRF.py:273-279: ``tf.image.resize(images, (224,224)) * 2 - 1`` (bilinear, half-pixel centres, no
    antialiasing).  images: (B,H,W,3) in [0,1] -> (B,3,224,224)."""
    x = images.to(torch.float32).permute(0, 3, 1, 2)
    x = F.interpolate(x, size=(EMBEDDER_INPUT_SIZE, EMBEDDER_INPUT_SIZE), mode="bilinear", align_corners=False,
                      antialias=False)
    return x * 2.0 - 1.0


def consistency_loss(embedding_source, embedding_target):
    """src/DietNeRF.py:261-270: ``(1 + keras.losses.cosine_similarity(s, t)) / 2``.  Keras' cosine_similarity is the
    NEGATIVE cosine (a loss), so this is (1 - cos(s,t)) / 2 in [0,1]: 0 when the embeddings align."""
    s = F.normalize(embedding_source.reshape(1, -1), dim=-1, eps=1e-12)
    t = F.normalize(embedding_target.reshape(1, -1), dim=-1, eps=1e-12)
    return ((1.0 - (s * t).sum()) / 2.0).squeeze()
