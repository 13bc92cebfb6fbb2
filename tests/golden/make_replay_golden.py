"""Golden windows for the replay batcher, from the LIVE reference (build container only).

    python tests/golden/make_replay_golden.py      # writes tests/golden/replay.npz

Runs the unmodified `tools.sample_episodes` / `tools.from_generator` of /root/reference on a
seeded synthetic episode store and saves the first batches.  tests/test_replay.py rebuilds the
same store from the same seed and checks the package's batcher against these bit for bit.
"""
import contextlib
import io
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))


def make_store(seed=0, n=7):
    """Episodes of ragged length (one shorter than 2, which the sampler must skip)."""
    rs = np.random.RandomState(seed)
    lens = [1] + [int(x) for x in rs.randint(3, 40, size=n - 1)]
    store = {}
    for i, T in enumerate(lens):
        first = np.zeros(T, bool); first[0] = True
        store["ep%d" % i] = {
            "vector": rs.randn(T, 5).astype(np.float32),
            "action": rs.uniform(-1, 1, (T, 3)).astype(np.float32),
            "reward": rs.randn(T).astype(np.float32),
            "is_first": first,
            "is_terminal": np.zeros(T, bool),
            "log_extra": rs.randn(T).astype(np.float32),
        }
    return store


def main():
    sys.path.insert(0, "/root/reference")
    import tools as ref_tools
    out = {}
    for name, (length, batch, seed) in {"a": (16, 4, 0), "b": (50, 3, 7)}.items():
        gen = ref_tools.from_generator(ref_tools.sample_episodes(make_store(), length, seed), batch)
        with contextlib.redirect_stdout(io.StringIO()):      # the reference prints a counter
            for step in range(3):
                for k, v in next(gen).items():
                    out["%s/%d/%s" % (name, step, k)] = v
    np.savez_compressed(os.path.join(HERE, "replay.npz"), **out)
    print("wrote", len(out), "arrays")


if __name__ == "__main__":
    main()
