"""Tap tables that express every dense convolution of the network as the gather-GEMM of dc_conv_desc.

    out[n, y, x, co] (+)= sum_t sum_ci in[n, y*s + dh[t], x*s + dw[t], ci] * W[wt[t]][ci][co]
    G[wt[t]][co][ci]  += sum_{n,y,x} in[n, y*s + dh[t], x*s + dw[t], ci] * dout[n, y, x, co]

Pure Python (no CUDA): the tables are validated on CPU against torch.nn.functional in tests/test_convdesc.py.
Weight-slice index is always kh*k + kw of the PyTorch parameter.
"""


def conv_fprop_taps(k, pad, dil):
    """nn.Conv2d forward (DX:145,149,60,74,291,426,430,434,360-366): in = x, out = y, stride = conv stride."""
    return [(kh * dil - pad, kw * dil - pad, kh * k + kw) for kh in range(k) for kw in range(k)]


def conv_dgrad_taps(k, pad, dil):
    """nn.Conv2d input gradient for stride 1: in = dy, out = dx, stride 1.
    dx[h, w] = sum dy[h + pad - kh*dil, w + pad - kw*dil] * W[:, :, kh, kw]"""
    return [(pad - kh * dil, pad - kw * dil, kh * k + kw) for kh in range(k) for kw in range(k)]


def conv_wgrad_taps(k, pad, dil):
    """nn.Conv2d weight gradient: in = x (gathered with the conv stride), dout = dy."""
    return conv_fprop_taps(k, pad, dil)


def convT_fprop_taps(k, stride, pad, ph, pw):
    """nn.ConvTranspose2d forward (DX:352,356,369,374) for output parity class (ph, pw):
    y[stride*i + ph, stride*j + pw] = sum x[i + dh, j + dw] * W[:, :, kh, kw] over taps with
    (ph + pad - kh) divisible by stride, dh = (ph + pad - kh) / stride.  in = x, out = y[:, ph::s, pw::s]."""
    taps = []
    for kh in range(k):
        if (ph + pad - kh) % stride:
            continue
        for kw in range(k):
            if (pw + pad - kw) % stride:
                continue
            taps.append(((ph + pad - kh) // stride, (pw + pad - kw) // stride, kh * k + kw))
    return taps


def convT_fused_fprop_taps():
    """nn.ConvTranspose2d(k3, s2, p1, op1) forward (DX:374) with all four output parity classes in ONE contraction:
    y[2i+a, 2j+b, co] = sum_{dh,dw in {0,1}} x[i+dh, j+dw] * W[:, co, a+1-2dh, b+1-2dw]   (kernel indices outside 0..2: zero)
    in = x, out channels = ((a*2+b), co), weight slice dh*2+dw of the DC_PACK_NTK_CONVT2 pack."""
    return [(dh, dw, dh * 2 + dw) for dh in range(2) for dw in range(2)]


def convT_dgrad_taps(k, pad):
    """nn.ConvTranspose2d input gradient: dx[i, j] = sum dy[s*i - pad + kh, s*j - pad + kw] * W[:, :, kh, kw];
    in = dy, out = dx, stride = s."""
    return [(kh - pad, kw - pad, kh * k + kw) for kh in range(k) for kw in range(k)]


def convT_wgrad_taps(k, pad):
    """nn.ConvTranspose2d weight gradient: in = dy (gathered with stride s), dout = x (enumerated);
    yields G[tap][ci_in][co_out], which is the parameter layout [ci_in][co_out][kh][kw] after unpacking."""
    return convT_dgrad_taps(k, pad)


def conv_out_size(h, k, stride, pad, dil):
    return (h + 2 * pad - dil * (k - 1) - 1) // stride + 1
