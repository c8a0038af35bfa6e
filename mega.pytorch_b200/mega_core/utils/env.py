"""utils/env.py of the reference: `setup_environment` is imported first by tools/*.py; nothing to set up here."""


def setup_environment():
    return None


setup_environment()
