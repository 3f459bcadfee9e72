"""CPU probe for the round-2 plan: a small-channel 3x3 convolution over [pixels][C] rows is a plain
3x3 convolution over SUPER-PIXELS (G consecutive pixels x C channels = 64 "channels") with Toeplitz-expanded weights,
i.e. exactly what conv_shift_kernel (stride 1) / conv_gather_kernel<STRIDE> (stride 2) already compute on the PL layout.

  level0: 16 -> 16, stride 1, G = 4:  x [B,16,H,W] == X [B,64,H,W/4];  W' [64,64,3,3]
  level1: 16 -> 32, stride 2, G = 4 in / 2... see below: outputs are grouped 2 px (stride 2 halves the width)

Checks the index algebra with torch.conv2d on random data; prints the density of the expanded weights
(the share of MACs that are not multiplications by zero)."""
import torch
import torch.nn.functional as F


def to_super(x, G):
    """[B,C,H,W] -> [B,G*C,H,W/G]: channel index = j*C + c for pixel j of the group."""
    B, C, H, W = x.shape
    return x.view(B, C, H, W // G, G).permute(0, 4, 1, 2, 3).reshape(B, G * C, H, W // G)


def from_super(X, G):
    B, GC, H, Ws = X.shape
    C = GC // G
    return X.view(B, G, C, H, Ws).permute(0, 2, 3, 4, 1).reshape(B, C, H, Ws * G)


def expand_weight(w, Gin, Gout, stride):
    """w [Co,Ci,3,3] -> W' [Gout*Co, Gin*Ci, 3, 3] over super-pixels.
    Output pixel j_out of output super-pixel X sits at x_out = Gout*X + j_out and reads input pixels
    x_in = stride*x_out + dx, dx in {-1,0,1}.  Input super-pixel index = stride_s*X + n with stride_s = stride*Gout/Gin,
    n in {-1,0,1}; pixel j_in of it sits at Gin*(stride_s*X + n) + j_in."""
    Co, Ci = w.shape[:2]
    assert (stride * Gout) % Gin == 0
    Wp = torch.zeros(Gout * Co, Gin * Ci, 3, 3, dtype=w.dtype)
    for j_out in range(Gout):
        for n in (-1, 0, 1):
            for j_in in range(Gin):
                dx = Gin * n + j_in - stride * j_out
                if dx in (-1, 0, 1):
                    Wp[j_out * Co:(j_out + 1) * Co, j_in * Ci:(j_in + 1) * Ci, :, n + 1] = w[:, :, :, dx + 1]
    return Wp


def check(Ci, Co, stride, Gin, Gout, H=12, W=16):
    torch.manual_seed(0)
    x = torch.randn(2, Ci, H, W, dtype=torch.float64)
    w = torch.randn(Co, Ci, 3, 3, dtype=torch.float64)
    ref = F.conv2d(x, w, stride=stride, padding=1)
    Wp = expand_weight(w, Gin, Gout, stride)
    stride_s = stride * Gout // Gin
    Y = F.conv2d(to_super(x, Gin), Wp, stride=(stride, stride_s), padding=1)      # zero border of ONE super-pixel
    got = from_super(Y, Gout)
    err = (got - ref).abs().max().item()
    dens = (Wp != 0).double().mean().item()
    print("Ci %d Co %d stride %d  Gin %d Gout %d -> super conv %d -> %d ch, col stride %d: max err %.1e, weight density %.2f"
          % (Ci, Co, stride, Gin, Gout, Gin * Ci, Gout * Co, stride_s, err, dens))
    assert err < 1e-12


if __name__ == "__main__":
    check(16, 16, 1, 4, 4)      # level0: 64 -> 64 over [H, W/4], a conv_shift_kernel shape
    check(16, 32, 2, 4, 2)      # level1: 64 -> 64 (2 px x 32 ch), row stride 2, super-column stride 1
    check(16, 32, 2, 4, 4, W=32)  # level1 alternative: 64 -> 128 (4 px x 32 ch), stride 2 in both: conv_gather<STRIDE>
    check(32, 64, 2, 2, 1)      # level2 entry from the (2 px x 32 ch) layout back to plain 64-channel rows
