#!/usr/bin/env python
"""Entry point with the reference's name and flags (train.py:311-333); see action_conditioned_gans_b200/train.py."""
from action_conditioned_gans_b200.train import main

if __name__ == "__main__":
    main()
