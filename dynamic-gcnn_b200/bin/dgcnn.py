#!/usr/bin/env python
"""Entry point, as /root/reference/bin/dgcnn.py:1-14:  dgcnn.py {train,inference,iotest} [flags]."""
import os
import sys

DGCNN_DIR = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, DGCNN_DIR)
from dgcnn import DGCNN_FLAGS  # noqa: E402


def main():
    flags = DGCNN_FLAGS()
    flags.parse_args()


if __name__ == "__main__":
    main()
