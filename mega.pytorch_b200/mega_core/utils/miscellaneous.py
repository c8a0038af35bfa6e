"""utils/miscellaneous.py: mkdir"""
import os


def mkdir(path):
    os.makedirs(path, exist_ok=True)
