"""List the variables of a TF V2 checkpoint (like TensorFlow's inspect_checkpoint) and check them against the name map this
package expects -- the first thing to run when a real `model.best` / `pwcnet.ckpt-595000` is available (DESIGN.md section 5:
the bundle reader and the `MaskNet//...` name map are not pinned against a TF-written file yet).
Usage: python tools/inspect_ckpt.py <prefix | prefix.index | prefix.data-00000-of-00001> [--check]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from unsupervised_detection_b200 import checkpoint as ck  # noqa: E402
from unsupervised_detection_b200.checkpoint import tf_names  # noqa: E402


def pwcnet_variable_names():
    """Names of the PWC-Net variables this package loads (models/PWCNet/model_pwcnet.py of the reference: featpyr conv{l}{a,aa,b},
    predict_flow conv{l}_{0..4} + flow{l}, ctxt dc_conv{l}{1..7}, upsample up_flow{l} / up_feat{l}; pyramid levels 6..2)."""
    names = []
    for l in range(1, 7):
        names += ['pwcnet/featpyr/conv%d%s' % (l, s) for s in ('a', 'aa', 'b')]
    for l in range(6, 1, -1):
        names += ['pwcnet/predict_flow/conv%d_%d' % (l, i) for i in range(5)] + ['pwcnet/predict_flow/flow%d' % l]
        names += ['pwcnet/ctxt/dc_conv%d%d' % (l, i) for i in range(1, 8)]
        if l > 2:
            names += ['pwcnet/upsample/up_flow%d' % l, 'pwcnet/upsample/up_feat%d' % l]
    return [n + suffix for n in names for suffix in ('/kernel', '/bias')]


def main(argv):
    prefix = ck.normalize_prefix(argv[1])
    rows = ck.list_variables(prefix)
    for name, shape, dtype in rows:
        print('%-70s %-18s %s' % (name, tuple(shape), getattr(dtype, '__name__', dtype)))
    print('%d variables, %d parameters' % (len(rows), sum(int(__import__("numpy").prod(s)) if s else 1 for _, s, _ in rows)))
    if '--check' in argv:
        have = set(n for n, _, _ in rows)
        from unsupervised_detection_b200 import params_init
        want = list(params_init.init_generator()) + list(params_init.init_recover()) + pwcnet_variable_names()
        for scope in ('MaskNet', 'FlownetS', 'pwcnet'):
            names = [k for k in want if k.startswith(scope + '/')]
            hit = {sep: sum(tf_names.to_tf_name(k, sep) in have for k in names) for sep in ('//', '/')}
            print('%-9s %3d variables expected; found with "//" spelling: %3d, with "/" spelling: %3d' % (scope, len(names), hit['//'], hit['/']))
            miss = [k for k in names if not any(tf_names.to_tf_name(k, s) in have for s in ('//', '/'))]
            for k in miss[:10]:
                print('   missing: %s  (tried %s)' % (k, [tf_names.to_tf_name(k, s) for s in ('//', '/')]))


if __name__ == '__main__':
    main(sys.argv)
