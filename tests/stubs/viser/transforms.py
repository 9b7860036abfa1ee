"""stub of viser.transforms (imported, never used, by reference main.py:8)"""
