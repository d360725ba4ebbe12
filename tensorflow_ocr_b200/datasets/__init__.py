"""Drop-in mirrors of the reference's ``datasets`` head functions."""
