#!/usr/bin/env python
"""The usage snippet of the reference's README (README.md:86-98: MaskedAffineAutoregressiveTransform(features=2,
hidden_features=4) + RandomPermutation, trained on a 2-D toy batch as in examples/toy_2d.py:57-68), run through this
package's drop-in API on a B200.  Only the import line differs from the reference.

    python examples/readme_maf_2d.py [--steps 300]
"""
import argparse
import math
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from flowconductor_b200 import distributions, flows, transforms  # reference: from flowcon import ...


def two_spirals(n, generator):
    """2-D toy data in the spirit of flowcon.datasets.load_plane_dataset("two_spirals") (synthetic, no download)."""
    t = torch.sqrt(torch.rand(n // 2, 1, generator=generator)) * 540 * (2 * math.pi) / 360
    x = torch.cat((-torch.cos(t) * t + torch.rand(n // 2, 1, generator=generator) * 0.5,
                   torch.sin(t) * t + torch.rand(n // 2, 1, generator=generator) * 0.5), 1)
    return torch.cat((x, -x)) / 3 + torch.randn(n // 2 * 2, 2, generator=generator) * 0.1


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=300)
    args = ap.parse_args()
    assert torch.cuda.is_available(), "this package runs on the GPU only"
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(0)
    # README.md:86-98
    base_dist = distributions.StandardNormal(shape=[2])
    transform = transforms.CompositeTransform([
        transforms.MaskedAffineAutoregressiveTransform(features=2, hidden_features=4),
        transforms.RandomPermutation(features=2),
    ])
    flow = flows.Flow(transform, base_dist).to(dev)
    optimizer = torch.optim.Adam(flow.parameters(), lr=1e-2)
    for step in range(args.steps):
        x = two_spirals(500, g).to(dev)           # examples/toy_2d.py:24: train minibatch 500
        optimizer.zero_grad()
        loss = -flow.log_prob(inputs=x).mean()
        loss.backward()
        optimizer.step()
        if step % 100 == 0 or step == args.steps - 1:
            print("step {:4d}  loss {:.4f}".format(step, loss.item()))
    with torch.no_grad():
        samples = flow.sample(10000)               # examples/toy_2d.py:49: 10 000 evaluation points
        lp = flow.log_prob(two_spirals(10000, g).to(dev))
    print("samples", tuple(samples.shape), "mean held-out log_prob {:.4f}".format(lp.mean().item()))


if __name__ == "__main__":
    main()
