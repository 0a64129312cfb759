"""Golden vectors for the image side of the driver, FROM THE REFERENCE (build container only).

    python tests/golden/make_golden_io.py        # writes tests/golden/io_my_img.npz

my_img.py cannot be imported (it parses argv and builds the model at import), so the two pure functions
`load_data` (my_img.py:47-71) and `my_transform` (my_img.py:73-89) are compiled from its AST where the file lies and
run on seeded random images; inputs (as uint8 arrays) and outputs are stored."""
import ast
import os

import numpy as np
import torch
from PIL import Image

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = "/root/reference/my_img.py"


def main():
    tree = ast.parse(open(SRC).read())
    keep = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name in ("load_data", "my_transform")]
    ns = {"np": np, "torch": torch, "Image": Image}
    exec(compile(ast.Module(body=keep, type_ignores=[]), SRC, "exec"), ns)
    out = {}
    rng = np.random.default_rng(7)
    for tag, (h, w) in {"small": (40, 60), "large": (56, 70)}.items():
        l = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
        r = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
        Image.fromarray(l).save("/tmp/_l.png")
        Image.fromarray(r).save("/tmp/_r.png")
        data = ns["load_data"]("/tmp/_l.png", "/tmp/_r.png")
        left, right, hh, ww = ns["my_transform"](data, 48, 64)
        out.update({f"{tag}:l": l, f"{tag}:r": r, f"{tag}:data": data, f"{tag}:left": left.numpy(),
                    f"{tag}:right": right.numpy(), f"{tag}:hw": np.array([hh, ww])})
    np.savez_compressed(os.path.join(HERE, "io_my_img.npz"), **out)
    print({k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
