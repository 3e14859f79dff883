"""stand-in: utils/parse.py:8,12 binds `user_error = gr.Error` at import"""


class Error(Exception):
    pass
