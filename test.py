#!/usr/bin/env python
"""Entry point with the reference's name and flags (test.py:15-47); see action_conditioned_gans_b200/test.py."""
from action_conditioned_gans_b200.test import main

if __name__ == "__main__":
    main()
