"""TEST INFRASTRUCTURE ONLY -- CPU fp32 restatement of OrienMaskYOLOFPNPlus.forward.

Not product code: only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu-baseline /
``--impl reference`` legs may import this.

A functional walk over a state dict carrying the reference's 524 key names, using the same
third-party arithmetic the reference uses (torch CPU conv2d = oneDNN, batch_norm, leaky_relu,
nearest interpolate).  Files followed:
  /root/reference/model/base.py:104-137,278-279   ConvBNRelu = conv(bias=False) + BN(eps 1e-5) + LeakyReLU(0.1)
  /root/reference/model/base.py:95-101            NearestUpsample
  /root/reference/model/backbone/darknet.py:6-15  residual block  x + conv3x3(conv1x1(x))
  /root/reference/model/backbone/darknet.py:18-54 DarkNet53 stages (1,2,8,8,4 blocks), outputs x32,x16,x8,x4
  /root/reference/model/orienmask_yolo_fpnplus.py:39-90  necks, routes, skips, heads, concat order, split

Pinned against the unmodified reference module (same state dict loaded with strict=True) by
tests/test_oracle.py when /root/reference is present, and against tests/golden/ otherwise.
"""
import torch
import torch.nn.functional as F

STAGE_BLOCKS = (1, 2, 8, 8, 4)          # darknet.py:21-25


def _cbl(sd, prefix, x, stride=1):
    """conv_bn_leaky at key prefix ``<prefix>.conv_block``."""
    w = sd[prefix + '.conv_block.0.weight']
    pad = w.shape[-1] // 2
    y = F.conv2d(x, w, None, stride=stride, padding=pad)
    y = F.batch_norm(y, sd[prefix + '.conv_block.1.running_mean'], sd[prefix + '.conv_block.1.running_var'],
                     sd[prefix + '.conv_block.1.weight'], sd[prefix + '.conv_block.1.bias'],
                     training=False, eps=1e-5)
    return F.leaky_relu(y, 0.1)


def _seq(sd, prefix, x, n):
    for i in range(n):
        x = _cbl(sd, '%s.%d' % (prefix, i), x)
    return x


def _stage(sd, name, x, n_blocks):
    x = _cbl(sd, name + '.0', x, stride=2)
    for b in range(1, n_blocks + 1):
        y = _cbl(sd, '%s.%d.conv.0' % (name, b), x)
        y = _cbl(sd, '%s.%d.conv.1' % (name, b), y)
        x = x + y
    return x


def _up(x, k):
    return F.interpolate(x, scale_factor=k, mode='nearest')


@torch.no_grad()
def forward_oracle(sd, x, return_features=False):
    """sd: state dict (reference key names, fp32 CPU); x: [B,3,H,W] fp32. Returns the 3x(bbox, orien) tuple.
    A state dict without skip convolutions (it has ``route8.0``) is the OrienMaskYOLO variant, model/orienmask_yolo.py:71-86."""
    sd = {k: v.float() for k, v in sd.items() if v.is_floating_point()}
    x = x.float()
    t = _cbl(sd, 'backbone.conv1', x)
    t = _stage(sd, 'backbone.conv2', t, STAGE_BLOCKS[0])
    x4 = _stage(sd, 'backbone.conv3', t, STAGE_BLOCKS[1])
    x8 = _stage(sd, 'backbone.conv4', x4, STAGE_BLOCKS[2])
    x16 = _stage(sd, 'backbone.conv5', x8, STAGE_BLOCKS[3])
    x32 = _stage(sd, 'backbone.conv6', x16, STAGE_BLOCKS[4])

    neck32 = _seq(sd, 'neck32', x32, 5)
    neck16 = _seq(sd, 'neck16', torch.cat([_up(_cbl(sd, 'route32.0', neck32), 2), x16], 1), 5)
    neck8 = _seq(sd, 'neck8', torch.cat([_up(_cbl(sd, 'route16.0', neck16), 2), x8], 1), 5)

    def bbox_head(name, f):
        f = _cbl(sd, name + '.0', f)
        return F.conv2d(f, sd[name + '.1.weight'], sd[name + '.1.bias'])

    bbox32 = bbox_head('bbox_head32', neck32)
    bbox16 = bbox_head('bbox_head16', neck16)
    bbox8 = bbox_head('bbox_head8', neck8)

    if 'route8.0.conv_block.0.weight' in sd:                       # model/orienmask_yolo.py:83
        cat4 = torch.cat([_up(_cbl(sd, 'route8.0', neck8), 2), x4], 1)
    else:                                                           # model/orienmask_yolo_fpnplus.py:85-86
        cat4 = torch.cat([_up(_cbl(sd, 'skip32.0', neck32), 8), _up(_cbl(sd, 'skip16.0', neck16), 4),
                          _up(_cbl(sd, 'skip8.0', neck8), 2), _cbl(sd, 'skip4', x4)], 1)
    o = _seq(sd, 'neck4', cat4, 5)
    o = _seq(sd, 'orien_head', o, 5)
    o = F.conv2d(o, sd['orien_head.5.weight'], sd['orien_head.5.bias'])
    nA2 = o.shape[1] // 3
    orien32, orien16, orien8 = torch.split(o, nA2, dim=1)
    out = ((bbox32, orien32.contiguous()), (bbox16, orien16.contiguous()), (bbox8, orien8.contiguous()))
    if return_features:
        return out, dict(x4=x4, x8=x8, x16=x16, x32=x32, neck32=neck32, neck16=neck16, neck8=neck8)
    return out
