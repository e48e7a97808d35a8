"""Entry script, as in the reference (mainShapley.py:1-3): python mainShapley.py [flags]"""
from shapleyserver.start import start

if __name__ == "__main__":
    start()
