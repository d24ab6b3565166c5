"""List the variables of a TF V2 checkpoint (like TensorFlow's inspect_checkpoint) and check them against the name map this
package expects -- the first thing to run when a real `model.best` / `pwcnet.ckpt-595000` is available (DESIGN.md section 5:
the bundle reader and the `MaskNet//...` name map are not pinned against a TF-written file yet).
Usage: python tools/inspect_ckpt.py <prefix | prefix.index | prefix.data-00000-of-00001> [--check]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from unsupervised_detection_b200 import checkpoint as ck  # noqa: E402
from unsupervised_detection_b200.checkpoint import tf_names  # noqa: E402


def main(argv):
    prefix = ck.normalize_prefix(argv[1])
    rows = ck.list_variables(prefix)
    for name, shape, dtype in rows:
        print('%-70s %-18s %s' % (name, tuple(shape), getattr(dtype, '__name__', dtype)))
    print('%d variables, %d parameters' % (len(rows), sum(int(__import__("numpy").prod(s)) if s else 1 for _, s, _ in rows)))
    if '--check' in argv:
        have = set(n for n, _, _ in rows)
        from oracle.params import make_params          # only for the list of internal names / shapes (a tool, not the product path)
        want = list(make_params(0).keys())
        for scope in ('MaskNet', 'FlownetS', 'pwcnet'):
            names = [k for k in want if k.startswith(scope + '/')]
            hit = {sep: sum(tf_names.to_tf_name(k, sep) in have for k in names) for sep in ('//', '/')}
            print('%-9s %3d variables expected; found with "//" spelling: %3d, with "/" spelling: %3d' % (scope, len(names), hit['//'], hit['/']))
            miss = [k for k in names if not any(tf_names.to_tf_name(k, s) in have for s in ('//', '/'))]
            for k in miss[:10]:
                print('   missing: %s  (tried %s)' % (k, [tf_names.to_tf_name(k, s) for s in ('//', '/')]))


if __name__ == '__main__':
    main(sys.argv)
