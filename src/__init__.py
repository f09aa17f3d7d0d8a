import os

PROJECT_DIR = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
