"""Layer table of the OrienMask DarkNet-53 + FPNPlus network (the only model on the hot path).

One place that knows the convolution list, their state-dict prefixes (the reference's 524 key
names) and the dataflow.  Reference structure being mirrored:
  /root/reference/model/backbone/darknet.py:18-54     stem + 5 stride-2 stages of 1/2/8/8/4 blocks
  /root/reference/model/orienmask_yolo_fpnplus.py:15-36,39-72   necks / routes / skips / heads
"""
from collections import namedtuple

# kind: 'cbl' = conv(bias=False)+BN+LeakyReLU(0.1) stored under <prefix>.conv_block.{0,1};
#       'conv' = plain conv with bias stored under <prefix>.{weight,bias}
ConvSpec = namedtuple('ConvSpec', 'prefix cin cout k stride kind')

STAGE_CHANNELS = (32, 64, 128, 256, 512)
STAGE_BLOCKS = (1, 2, 8, 8, 4)


def conv_specs(num_anchors=3, num_classes=80, plus=True):
    """Every convolution of the network in execution order.  plus=False: OrienMaskYOLO (model/orienmask_yolo.py:8-86), the
    variant without the skip convolutions whose neck4 reads cat[up2(route8(neck8)), x4] (192 channels)."""
    specs = []

    def cbl(prefix, cin, cout, k, stride=1):
        specs.append(ConvSpec(prefix, cin, cout, k, stride, 'cbl'))

    def neck(prefix, cin, c):
        cbl(prefix + '.0', cin, c, 1)
        cbl(prefix + '.1', c, 2 * c, 3)
        cbl(prefix + '.2', 2 * c, c, 1)
        cbl(prefix + '.3', c, 2 * c, 3)
        cbl(prefix + '.4', 2 * c, c, 1)

    cbl('backbone.conv1', 3, 32, 3)
    for i, (c, n) in enumerate(zip(STAGE_CHANNELS, STAGE_BLOCKS)):
        stage = 'backbone.conv%d' % (i + 2)
        cbl(stage + '.0', c, 2 * c, 3, 2)
        for b in range(1, n + 1):
            cbl('%s.%d.conv.0' % (stage, b), 2 * c, c, 1)
            cbl('%s.%d.conv.1' % (stage, b), c, 2 * c, 3)
    neck('neck32', 1024, 512)
    cbl('route32.0', 512, 256, 1)
    neck('neck16', 768, 256)
    cbl('route16.0', 256, 128, 1)
    neck('neck8', 384, 128)
    bbox_dim = num_anchors * (5 + num_classes)
    for s, c in ((32, 512), (16, 256), (8, 128)):
        cbl('bbox_head%d.0' % s, c, 2 * c, 3)
        specs.append(ConvSpec('bbox_head%d.1' % s, 2 * c, bbox_dim, 1, 1, 'conv'))
    if plus:
        cbl('skip32.0', 512, 64, 1)
        cbl('skip16.0', 256, 64, 1)
        cbl('skip8.0', 128, 64, 1)
        cbl('skip4', 128, 64, 1)
        neck('neck4', 256, 128)
    else:
        cbl('route8.0', 128, 64, 1)
        neck('neck4', 192, 128)
    for i, (cin, cout, k) in enumerate(((128, 256, 3), (256, 128, 1), (128, 256, 3), (256, 128, 1), (128, 256, 3))):
        cbl('orien_head.%d' % i, cin, cout, k)
    specs.append(ConvSpec('orien_head.5', 256, num_anchors * 6, 1, 1, 'conv'))
    return specs


def state_dict_shapes(num_anchors=3, num_classes=80, plus=True):
    """Ordered {key: shape} of the reference state dict (524 entries for 3 anchors / 80 classes)."""
    out = {}
    for s in conv_specs(num_anchors, num_classes, plus):
        if s.kind == 'cbl':
            out[s.prefix + '.conv_block.0.weight'] = (s.cout, s.cin, s.k, s.k)
            for name in ('weight', 'bias', 'running_mean', 'running_var'):
                out['%s.conv_block.1.%s' % (s.prefix, name)] = (s.cout,)
            out[s.prefix + '.conv_block.1.num_batches_tracked'] = ()
        else:
            out[s.prefix + '.weight'] = (s.cout, s.cin, s.k, s.k)
            out[s.prefix + '.bias'] = (s.cout,)
    return out


def macs_per_image(height, width, num_anchors=3, num_classes=80, plus=True):
    """Multiply-accumulates of the 90 convolutions for one HxW image (86.9226e9 @544)."""
    res = {}
    total = 0
    h, w = height, width
    # resolution of every conv output, derived from the dataflow
    strides = {}
    cur = 1
    for s in conv_specs(num_anchors, num_classes, plus):
        p = s.prefix
        if p.startswith('backbone'):
            cur *= s.stride
            strides[p] = cur
        elif p.startswith(('neck32', 'route32', 'bbox_head32', 'skip32')):
            strides[p] = 32
        elif p.startswith(('neck16', 'route16', 'bbox_head16', 'skip16')):
            strides[p] = 16
        elif p.startswith(('neck8', 'bbox_head8', 'skip8', 'route8')):
            strides[p] = 8
        else:
            strides[p] = 4
        st = strides[p]
        m = (h // st) * (w // st) * s.cout * s.cin * s.k * s.k
        res[p] = m
        total += m
    return total, res
