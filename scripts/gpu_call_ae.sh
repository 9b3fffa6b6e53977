#!/bin/bash
timeout 400 python -m pytest tests/ -q -m gpu --timeout 120 2>&1 | tail -6
