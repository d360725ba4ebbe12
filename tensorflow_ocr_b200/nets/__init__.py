"""Drop-in mirrors of the reference's ``nets`` head functions."""
