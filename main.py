#!/usr/bin/env python3
"""Entry point with the reference's surface (reference main.py:28-31):

    python main.py predict --input data/test --output results --model models/best_model.pth
"""
import sys

from unet_watermark_b200.cli import main

if __name__ == "__main__":
    sys.exit(main())
