from dca_b200.__main__ import main, parse_args  # noqa: F401

if __name__ == '__main__':
    main()
