"""Emit the DeeperCut deploy net as Caffe text-format prototxt.

The reference ships ``models/deepercut/ResNet-152.prototxt`` (7344 lines); the
GPU box has no /root/reference, and reference files are not copied into this
repo, so the same NetParameter is *generated*: bottleneck counts per stage
(3, 8, 36, 3) = ResNet-152 as shipped; (3, 4, 23, 3) = the ResNet-101 variant
SURVEY.md section 0 describes.  ``tests/test_prototxt.py`` checks (here, where the
reference is mounted) that the generated text parses to a structure identical
to the reference file's, layer for layer and field for field.

Topology facts follow ResNet-152.prototxt: conv1 7x7/2 (:12-24), pool1 MAX 3x3/2
(:56-66), res5 stride removed + dilation 2 (:6760-6770, :6847), heads (:7219-7345).
"""
import sys

STAGES_152 = (3, 8, 36, 3)
STAGES_101 = (3, 4, 23, 3)


def _conv(name, bottom, top, nout, k, pad, stride, bias_term=False, dilation=None):
    s = ['layer {', '  bottom: "%s"' % bottom, '  top: "%s"' % top, '  name: "%s"' % name,
         '  type: "Convolution"', '  convolution_param {', '    num_output: %d' % nout,
         '    kernel_size: %d' % k]
    if dilation:
        s.append('    dilation: %d' % dilation)
    s += ['    pad: %d' % pad, '    stride: %d' % stride]
    if not bias_term:
        s.append('    bias_term: false')
    s += ['  }', '}']
    return s


def _bn_scale(blob, suffix):
    return ['layer {', '  bottom: "%s"' % blob, '  top: "%s"' % blob, '  name: "bn%s"' % suffix,
            '  type: "BatchNorm"', '  param { lr_mult: 0 }', '  param { lr_mult: 0 }',
            '  param { lr_mult: 0 }', '  batch_norm_param { use_global_stats: true }', '}',
            'layer {', '  bottom: "%s"' % blob, '  top: "%s"' % blob, '  name: "scale%s"' % suffix,
            '  type: "Scale"', '  scale_param { bias_term: true }', '}']


def _relu(blob, name):
    return ['layer {', '  top: "%s"' % blob, '  bottom: "%s"' % blob, '  name: "%s"' % name,
            '  type: "ReLU"', '}']


def _block_names(stage_idx, nblocks):
    """res2a res2b res2c | res3a res3b1.. | res4a res4b1.. | res5a res5b res5c."""
    stage = stage_idx + 2
    if nblocks == 3:
        return ["%d%s" % (stage, c) for c in "abc"]
    return ["%da" % stage] + ["%db%d" % (stage, i) for i in range(1, nblocks)]


def generate(stages=STAGES_152, height=688, width=688, name=None, heads=(("pose", 14), ("locref", 28), ("next", 364))):
    L = ['name: "%s"' % (name or ("ResNet-152" if tuple(stages) == STAGES_152 else "ResNet-101")),
         'input: "data"', 'input_dim: 1', 'input_dim: 3', 'input_dim: %d' % height, 'input_dim: %d' % width]
    L += _conv("conv1", "data", "conv1", 64, 7, 3, 2)
    L += _bn_scale("conv1", "_conv1") + _relu("conv1", "conv1_relu")
    L += ['layer {', '  bottom: "conv1"', '  top: "pool1"', '  name: "pool1"', '  type: "Pooling"',
          '  pooling_param { kernel_size: 3 stride: 2 pool: MAX }', '}']
    prev = "pool1"
    skip_tap = None
    for si, nb in enumerate(stages):
        mid = 64 << si
        out = 256 << si
        for bi, bn in enumerate(_block_names(si, nb)):
            first = bi == 0
            # stride 2 lives in the 1x1 branch1/branch2a of res3a/res4a; res5a keeps stride 1
            stride = 2 if (first and si in (1, 2)) else 1
            dil = 2 if si == 3 else None
            if first:
                b1 = "res%s_branch1" % bn
                L += _conv(b1, prev, b1, out, 1, 0, stride)
                L += _bn_scale(b1, "%s_branch1" % bn)
                shortcut = b1
            else:
                shortcut = prev
            a, b, c = ("res%s_branch2%s" % (bn, x) for x in "abc")
            L += _conv(a, prev, a, mid, 1, 0, stride)
            L += _bn_scale(a, "%s_branch2a" % bn) + _relu(a, "res%s_branch2a_relu" % bn)
            L += _conv(b, a, b, mid, 3, dil or 1, 1, dilation=dil)
            L += _bn_scale(b, "%s_branch2b" % bn) + _relu(b, "res%s_branch2b_relu" % bn)
            L += _conv(c, b, c, out, 1, 0, 1)
            L += _bn_scale(c, "%s_branch2c" % bn)
            res = "res%s" % bn
            L += ['layer {', '  bottom: "%s"' % shortcut, '  bottom: "%s"' % c, '  top: "%s"' % res,
                  '  name: "%s"' % res, '  type: "Eltwise"', '}']
            L += _relu(res, "%s_relu" % res)
            prev = res
        if si == 1:
            skip_tap = prev
    crop_names = {"pose": "crop1", "locref": "crop_locref", "next": "crop_next"}
    out_names = {"pose": "fc_pose", "locref": "loc_pred", "next": "next_pred"}
    for head, nout in heads:
        up, sk = "%s_up_%s" % (prev, head), "res3d_%s" % head
        L += ['layer {', '  bottom: "%s"' % prev, '  top: "%s"' % up, '  name: "%s"' % up,
              '  type: "Deconvolution"', '  convolution_param { num_output: %d kernel_size: 3 pad: 0 stride: 2 }' % nout, '}']
        L += _conv(sk, skip_tap, sk, nout, 1, 0, 1, bias_term=True)
        L += ['layer { type: "Crop" name: "%s" bottom: "%s" bottom: "%s" top: \'%sc\' }'
              % (crop_names.get(head, "crop_" + head), up, sk, up)]
        L += ['layer {', '  bottom: "%s"' % sk, '  bottom: "%sc"' % up, '  top: "%s"' % out_names.get(head, head),
              '  name: "%s"' % out_names.get(head, head), '  type: "Eltwise"', '}']
        if head == "pose":
            L += ['layer {', '  name: "prob"', '  type: "Sigmoid"', '  bottom: "fc_pose"', '  top: "prob"', '}']
    return "\n".join(L) + "\n"


def write(path, **kw):
    with open(path, "w") as f:
        f.write(generate(**kw))
    return path


if __name__ == "__main__":
    out = sys.argv[1] if len(sys.argv) > 1 else "/dev/stdout"
    stages = STAGES_101 if (len(sys.argv) > 2 and sys.argv[2] == "101") else STAGES_152
    write(out, stages=stages)
