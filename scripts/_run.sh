python -m pytest tests -m gpu -x -q -s -k "trigonometry" 2>&1 | grep -E "bit-equal|passed|failed|assert|Error" | tail -6
