"""Phase timeline of the persistent IndexTTS-GPT decode kernel from a B200TTS_GPT_TRACE dump ([cta][layer][16] globaltimer ns).
    python tools/igpt_trace.py gpurun_out/igpt_trace.bin [layer ...]"""
import sys
import numpy as np
t = np.fromfile(sys.argv[1], dtype=np.uint64)
L = t.size // (148 * 16)
t = t.reshape(148, L, 16).astype(np.int64)
rel = (t - t[:, :, 0][t[:, :, 0] > 0].min()) / 1000.0
rel[t == 0] = np.nan
names = ['P1 start', 'P1 polled', 'P1 ln', 'P1 gemv', 'att done', 'P3 start', 'P3 polled', 'P4 start', 'P4 polled', 'P4 ln', 'P5 start',
         'P5 polled', 'layer end']
print('mean layer duration (us), CTA 0:', round(float(np.nanmean(np.diff(rel[0, :, 0]))), 2))
for l in [int(x) for x in sys.argv[2:]] or [L // 2]:
    base = np.nanmin(rel[:, l, 0])
    print('layer', l)
    for i, n in enumerate(names):
        col = rel[:, l, i] - base
        print('  %-10s min %6.2f mean %6.2f max %6.2f  ctas=%d' % (n, np.nanmin(col), np.nanmean(col), np.nanmax(col), np.sum(~np.isnan(col))))
