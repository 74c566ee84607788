#!/bin/bash
# whole GPU suite
python -m pytest tests -x -q -m gpu -s 2>&1 | grep -v "^$" | tail -25
