"""TEST INFRASTRUCTURE ONLY -- CPU restatement (PyTorch fp32, eager) of the reference's algorithm
for the hot path. Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this package; the product (text-to-speech-tts-onnx_b200/) never does.

Parity pinning: the reference ships no golden vectors or tests for this path, so
the restatement is pinned against the reference's own nn.Module sources executed in the build
container (oracle/ref_harness.py imports them from /root/reference through small stubs) and the
resulting vectors are committed under tests/golden/ (made by oracle/make_golden.py)."""
