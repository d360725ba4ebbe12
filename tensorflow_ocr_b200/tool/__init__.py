"""Drop-in mirrors of the reference's ``tool`` head functions."""
