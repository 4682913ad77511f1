"""Digest of a DPX_TRACE dump (phase timestamps of k_col CTAs / persistent row-kernel tiles)."""
import sys
import numpy as np
for name in sys.argv[1:]:
    a = np.fromfile(name, dtype=np.uint64).reshape(2, 16384, 16)
    for k, label in ((0, 'col'), (1, 'row')):
        t = a[k]
        t = t[t[:, 1] > 0].astype(np.int64)
        if len(t) == 0:
            continue
        t0 = t[:, 1].min()
        slots = [i for i in range(1, 8) if (t[:, i] > 0).all()]
        ph = t[:, slots] - t0
        d = np.diff(ph, axis=1)
        print(name, label, 'records', len(t), 'span us %.1f' % ((t[:, 7].max() - t0) / 1e3))
        print('  slots', slots, 'phase us mean', np.round(d.mean(0) / 1e3, 2), 'total', round((ph[:, -1] - ph[:, 0]).mean() / 1e3, 2))
