"""Re-record the host launch sequences checked by tests/test_host_sequence.py (CPU dry run, nothing is computed).

    python tests/golden/make_host_sequences.py

Run it only after the GPU parity suite (`pytest -m gpu`) has been seen green with the engine being recorded: the
fixtures then pin the kernel sequence and buffer wiring of THAT engine against host refactors that cannot be run on a GPU.
"""
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, os.path.dirname(HERE))

import _dryrun  # noqa: E402


def main():
    import dcanet_b200 as d
    tr, _ = _dryrun.forward_trace(d.GwcNet(192).eval(), 96, 312)
    json.dump(tr, open(os.path.join(HERE, "kernel_sequence_kitti.json"), "w"))
    tr2, _ = _dryrun.forward_trace(d.GwcNet(48).eval(), 16, 32)
    json.dump(tr2, open(os.path.join(HERE, "kernel_sequence_tiny.json"), "w"))
    json.dump(_dryrun.dataflow_trace(d.GwcNet(48).eval(), 16, 32), open(os.path.join(HERE, "kernel_dataflow_tiny.json"), "w"))
    print("launches per KITTI forward:", len(tr), "; tiny:", len(tr2))


if __name__ == "__main__":
    main()
